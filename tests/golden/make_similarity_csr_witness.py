#!/usr/bin/env python3
"""Independent witnesses for the two other results the reference holds no golden vector for -- the integer part of
Similarity::set_table (reference src/analyses/similarity.rs:125-150) and the AbacusByGroup CSR {r, c, v}
(src/graph_broker/abacus.rs:859-986) -- by brute force straight from the S / P lines of the fixture GFAs: no ItemTable,
no cursor passes, nothing shared with oracle/ or the C++ host layer.  Writes tests/golden/similarity_csr_witness.json.

    inter[x][y] = sum over items of w * [item in group x] * [item in group y]        (similarity.rs:138-146)
    len[x]      = sum over items of w * [item in group x]                            (similarity.rs:131-136)
    w = 1 (node, edge) or the segment length (bp)
    CSR (node / bp tables only; items = segments in S-line order, ids from 1, row 0 = the dummy item):
      r[i + 1] - r[i] = number of distinct groups containing item i, c = those groups in ascending order,
      v = how many steps of the group's paths visit the item                         (abacus.rs:790-799)
Group index = rank of first appearance of the group among the P lines (abacus.rs:310-347).
"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_ordered_growth_witness import group_of, read_gfa  # the (oracle-independent) GFA reader of the first witness


def segment_order(path):
    return [line.split("\t")[1] for line in open(path) if line.startswith("S\t")]


def witness(gfa, mode, count):
    seg_len, paths = read_gfa(gfa)
    groups = []
    for name, _ in paths:
        g = group_of(name, mode)
        if g not in groups:
            groups.append(g)
    G = len(groups)
    occ = {}  # item -> {group index: number of visits}
    weight = {}
    for name, steps in paths:
        gi = groups.index(group_of(name, mode))
        if count in ("node", "bp"):
            for node, _ in steps:
                occ.setdefault(node, {}).setdefault(gi, 0)
                occ[node][gi] += 1
                weight[node] = 1 if count == "node" else seg_len[node]
        else:
            for (u, fu), (v, fv) in zip(steps, steps[1:]):
                a, b = (u, fu, v, fv), (v, not fv, u, not fu)
                e = a if a <= b else b
                occ.setdefault(e, {}).setdefault(gi, 0)
                occ[e][gi] += 1
                weight[e] = 1
    inter = [[0] * G for _ in range(G)]
    length = [0] * G
    for item, per in occ.items():
        gs = sorted(per)
        for x in gs:
            length[x] += weight[item]
            for y in gs:
                inter[x][y] += weight[item]
    out = {"gfa": os.path.basename(gfa), "grouping": mode, "count": count, "groups": groups, "inter": inter, "len": length}
    if count in ("node", "bp"):
        r, c, v = [0, 0], [], []  # row 0: the dummy item has no entries
        for seg in segment_order(gfa):
            per = occ.get(seg, {})
            for g in sorted(per):
                c.append(g)
                v.append(per[g])
            r.append(len(c))
        out.update({"r": r, "c": c, "v": v})
    return out


def main():
    cases = []
    for gfa, modes in (("chrM_test.gfa", ["sample", "haplotype", "path"]), ("t_groups.gfa", ["path", "sample", "haplotype"]), ("cdbg.gfa", ["path"])):
        for mode in modes:
            for count in ("node", "bp", "edge"):
                cases.append(witness(os.path.join(HERE, gfa), mode, count))
    json.dump({"source": "brute force over the S / P lines (tests/golden/make_similarity_csr_witness.py)", "cases": cases},
              open(os.path.join(HERE, "similarity_csr_witness.json"), "w"))
    print(len(cases), "cases written")


if __name__ == "__main__":
    sys.exit(main())
