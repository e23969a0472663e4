#!/usr/bin/env python3
"""Independent witness for ordered growth: brute-force evaluation of the rule of AbacusByGroup::calc_growth
(reference src/graph_broker/abacus.rs:1003-1010) DIRECTLY from the P lines of the fixture GFAs -- no ItemTable, no CSR,
nothing shared with oracle/ or the C++ host layer.  The reference holds no golden vector for ordered growth; this
script writes tests/golden/ordered_growth_witness.json, which pins both the oracle and the CUDA path.

For an item present in the groups g_1 < g_2 < ... < g_m (group index = rank of first appearance of the group among the
P lines, all paths of a group together, abacus.rs:310-347) and thresholds (c, q), C = max(1, c):
    the item is skipped unless m >= C                                                            (abacus.rs:1003)
    for every column j >= g_1: let g_k be the last of its groups with g_k <= j;
    it adds its weight to res[j] iff k >= ceil((g_k + 1) * q)        (abacus.rs:1005-1010, k counted from 1)
weight = 1 (node, edge) or the segment's sequence length (bp).  Edges: consecutive steps of a path, canonical
orientation as in graph.rs:142-148, identified by the pair of oriented nodes.
"""
import json
import math
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))


def read_gfa(path):
    seg_len, paths = {}, []
    for line in open(path):
        f = line.rstrip("\n").split("\t")
        if f[0] == "S":
            seg_len[f[1]] = len(f[2])
        elif f[0] == "P":
            steps = [(s[:-1], s[-1] == "+") for s in f[2].split(",") if s]
            paths.append((f[1], steps))
    return seg_len, paths


def group_of(name, mode):
    if mode == "path":
        return name
    parts = name.split("#")
    if mode == "sample":
        return parts[0]
    return parts[0] + "#" + (parts[1] if len(parts) > 1 else "")  # label as in abacus.rs:248-262 ("sample#" when no haplotype)


def witness(gfa, mode, count, pairs):
    seg_len, paths = read_gfa(gfa)
    groups = []
    for name, _ in paths:
        g = group_of(name, mode)
        if g not in groups:
            groups.append(g)
    G = len(groups)
    member = {}   # item -> set of group indices
    weight = {}
    for name, steps in paths:
        gi = groups.index(group_of(name, mode))
        if count in ("node", "bp"):
            for node, _ in steps:
                member.setdefault(node, set()).add(gi)
                weight[node] = 1 if count == "node" else seg_len[node]
        else:
            for (u, fu), (v, fv) in zip(steps, steps[1:]):
                a, b = (u, fu, v, fv), (v, not fv, u, not fu)
                e = a if a <= b else b   # any fixed choice between the two orientations identifies the edge
                member.setdefault(e, set()).add(gi)
                weight[e] = 1
    curves = []
    for c, q in pairs:
        C = max(1, c)
        res = [0] * G
        for item, gs in member.items():
            g = sorted(gs)
            if len(g) < C:
                continue
            for j in range(g[0], G):
                k = max(i for i, x in enumerate(g) if x <= j)          # last group of the item at or before column j
                if k + 1 >= math.ceil((g[k] + 1.0) * max(0.0, q)):
                    res[j] += weight[item]
        curves.append(res)
    return groups, curves


def main():
    pairs = [(1, 0.0), (2, 0.0), (1, 0.5), (2, 0.5), (1, 1.0), (3, 0.9), (0, 0.1), (2, 0.33)]
    cases = []
    for gfa, modes in (("chrM_test.gfa", ["sample", "haplotype", "path"]), ("t_groups.gfa", ["path", "sample", "haplotype"]), ("cdbg.gfa", ["path"])):
        for mode in modes:
            for count in ("node", "bp", "edge"):
                groups, curves = witness(os.path.join(HERE, gfa), mode, count, pairs)
                cases.append({"gfa": gfa, "grouping": mode, "count": count, "groups": groups, "curves": curves})
    out = {"source": "brute force over the P lines, rule of src/graph_broker/abacus.rs:1003-1010 (tests/golden/make_ordered_growth_witness.py)",
           "pairs": [list(p) for p in pairs], "cases": cases}
    json.dump(out, open(os.path.join(HERE, "ordered_growth_witness.json"), "w"))
    print(len(cases), "cases written")


if __name__ == "__main__":
    sys.exit(main())
