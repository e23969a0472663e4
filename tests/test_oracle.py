"""Pins the CPU oracle (oracle/) against every golden vector the reference holds for this path
(SURVEY.md section 8c).  CPU only."""
import json
import math
import os

import numpy as np
import pytest

from oracle import gfa_oracle as go
from oracle import oracle as po

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
KATS = json.load(open(os.path.join(GOLDEN, "kats.json")))


# ---- closed-form growth KATs: src/graph_broker/hist.rs:341-398 (exact f64 equality) ---------------

def test_choose_kat():
    # hist.rs:341-349
    assert abs(po.choose(5, 0) - 0.0) < 1e-10
    assert abs(po.choose(5, 1) - math.log2(5.0)) < 1e-10
    assert abs(po.choose(5, 5) - 0.0) < 1e-10
    assert abs(po.choose(5, 2) - math.log2(10.0)) < 1e-10


def test_growth_union_kat():
    k = KATS["union"]
    assert list(po.growth_union(k["hist"], po.absolute(k["coverage"]))) == k["expect"]


def test_growth_core_kat():
    k = KATS["core"]
    assert list(po.growth_core(k["hist"], po.absolute(k["coverage"]))) == k["expect"]


def test_growth_quorum_kat():
    k = KATS["quorum"]
    got = po.growth_quorum(k["hist"], po.absolute(k["coverage"]), po.relative(k["quorum"]))
    assert list(got) == k["expect"]


def test_chr22_growth_golden():
    """docs/chr22.hprc-v1.0-pggb.histgrowth.html:267-274: 3 count types x 5 threshold pairs x 44 points."""
    d = json.load(open(os.path.join(GOLDEN, "chr22_histgrowth.json")))
    checked = 0
    for count in ("bp", "node", "edge"):
        hist = d["hist"][count]["values"]
        g = d["growth"][count]
        for cov, quo, curve in zip(g["coverage"], g["quorum"], g["curves"]):
            got = po.hist_calc_growth(hist, po.absolute(int(cov)), po.relative(float(quo)))
            assert [math.floor(x) for x in got] == [int(v) for v in curve], (count, cov, quo)
            checked += len(curve)
    assert checked == 3 * 5 * 44


# ---- coverage / hist goldens: abacus.rs:1485-1633, tests/test_files/t_groups.hist.tsv ------------

def _tables(gfa, count, **mask_kw):
    g = go.parse_gfa(os.path.join(GOLDEN, gfa))
    mask = go.make_mask(g, **mask_kw)
    t = go.item_tables(g, mask, count)
    op, og, names = go.path_order_arrays(mask, g)
    return g, t, op, og, names


@pytest.mark.parametrize("count", ["node", "edge", "bp"])
def test_chrM_hist_golden(count):
    g, t, op, og, names = _tables("chrM_test.gfa", count, groupby_sample=True)
    k = KATS["chrM_groupby_sample"]
    assert names == k["groups"]
    countable = po.abacus_by_total(t.n_items, t.items, t.id_prefsum, op, og, t.exclude)
    assert countable[0] == 0xFFFFFFFF
    assert list(countable[1:]) == k["countable_" + count]
    if count == "bp":
        hist = po.construct_hist_bps(countable, g.node_lens, len(names), t.uncovered)
    else:
        hist = po.construct_hist(countable, len(names))
    assert list(hist) == k[count]


def test_t_groups_hist_golden():
    g, t, op, og, names = _tables("t_groups.gfa", "node")
    countable = po.abacus_by_total(t.n_items, t.items, t.id_prefsum, op, og, t.exclude)
    hist = po.construct_hist(countable, len(names))
    assert list(hist) == KATS["t_groups_node_hist"]
    rows = [l.split("\t") for l in open(os.path.join(GOLDEN, "t_groups.hist.tsv")).read().splitlines()
            if l and l[0].isdigit()]
    assert [int(r[1]) for r in rows] == [int(x) for x in hist]


# ---- ordered growth: no reference golden; two independent formulations must agree -------------------

def test_ordered_growth_known_answers():
    """Derived values recorded in SURVEY.md section 4 (secondary goldens)."""
    g, t, op, og, names = _tables("chrM_test.gfa", "node", groupby_sample=True)
    r, c, v = po.csr_build(t.n_items, t.items, t.id_prefsum, op, og, t.exclude)
    assert list(po.calc_growth(r, c, 4, po.absolute(1), po.relative(0.0))) == [89, 106, 140, 154]
    assert list(po.calc_growth(r, c, 4, po.absolute(2), po.relative(0.5))) == [87, 101, 101, 115]
    bp = po.calc_growth(r, c, 4, po.absolute(1), po.relative(0.0), count_bp=True, node_lens=g.node_lens)
    assert list(bp) == [16569, 17147, 17183, 17197]
    g, t, op, og, names = _tables("t_groups.gfa", "node")
    r, c, v = po.csr_build(t.n_items, t.items, t.id_prefsum, op, og, t.exclude)
    assert names == ["y#1", "y#2", "y#3", "y#4", "y#5", "x"]
    assert list(po.calc_growth(r, c, 6, po.absolute(1), po.relative(0.0))) == [2, 5, 8, 9, 10, 10]


def test_ordered_growth_end_point_is_pinned_by_hist():
    """curve[-1] for c=1,q=0 equals sum(hist[1:]) -- ties ordered growth to the pinned hists."""
    for count in ("node", "edge"):
        g, t, op, og, names = _tables("chrM_test.gfa", count, groupby_sample=True)
        r, c, v = po.csr_build(t.n_items, t.items, t.id_prefsum, op, og, t.exclude)
        curve = po.calc_growth(r, c, len(names), po.absolute(1), po.relative(0.0))
        assert curve[-1] == sum(KATS["chrM_groupby_sample"][count][1:])


@pytest.mark.parametrize("seed", range(6))
def test_csr_walk_equals_bitmap_rule(seed):
    rng = np.random.default_rng(seed)
    N, G = int(rng.integers(1, 300)), int(rng.integers(1, 200))
    dens = rng.choice([0.02, 0.3, 0.6, 0.95])
    bits = (rng.random((N + 1, G)) < dens).astype(np.uint8)
    bits[0] = 0
    W = (G + 63) // 64
    padded = np.zeros((N + 1, W * 64), dtype=np.uint8)
    padded[:, :G] = bits
    bitmap = np.packbits(padded, axis=1, bitorder="little").view(np.uint64).reshape(N + 1, W)
    weights = rng.integers(0, 1000, N + 1).astype(np.uint32)
    items, prefsum, op, og = po.bitmap_to_item_table(bitmap, G)
    r, c, v = po.csr_build(N, items, prefsum, op, og)
    assert int(r[-1]) == int(bits.sum())
    for cov in (1, 2, max(1, G // 2)):
        for q in (0.0, 0.1, 0.5, 0.9, 1.0):
            ref = po.calc_growth(r, c, G, po.absolute(cov), po.relative(q))
            alt = po.ordered_growth_bitmap_rule(bitmap, G, cov, q)
            assert [int(x) for x in ref] == [int(x) for x in alt], (cov, q)
            refw = po.calc_growth(r, c, G, po.absolute(cov), po.relative(q), count_bp=True, node_lens=weights)
            altw = po.ordered_growth_bitmap_rule(bitmap, G, cov, q, weights)
            assert [int(x) for x in refw] == [int(x) for x in altw], (cov, q)


def test_csr_is_plain_sorted_csr():
    rng = np.random.default_rng(7)
    N, G = 120, 37
    bits = (rng.random((N + 1, G)) < 0.4).astype(np.uint8)
    bits[0] = 0
    bitmap = np.packbits(np.pad(bits, ((0, 0), (0, 64 - G))), axis=1, bitorder="little").view(np.uint64).reshape(N + 1, 1)
    items, prefsum, op, og = po.bitmap_to_item_table(bitmap, G)
    # visit every item twice in every path: v counts occurrences, c/r unchanged
    items2 = np.concatenate([np.repeat(items[int(prefsum[g]):int(prefsum[g + 1])], 2) for g in range(G)])
    prefsum2 = prefsum * np.uint64(2)
    r, c, v = po.csr_build(N, items2, prefsum2, op, og)
    for i in range(1, N + 1):
        row = c[int(r[i]):int(r[i + 1])]
        assert list(row) == list(np.nonzero(bits[i])[0])
        assert all(int(x) == 2 for x in v[int(r[i]):int(r[i + 1])])


def test_similarity_matches_dense_algebra():
    rng = np.random.default_rng(3)
    N, G = 200, 23
    bits = (rng.random((N + 1, G)) < 0.5).astype(np.uint8)
    bits[0] = 0
    bits[1, :] = 1  # every group non-empty
    bitmap = np.packbits(np.pad(bits, ((0, 0), (0, 64 - G))), axis=1, bitorder="little").view(np.uint64).reshape(N + 1, 1)
    w = rng.integers(1, 50, N + 1).astype(np.uint32)
    items, prefsum, op, og = po.bitmap_to_item_table(bitmap, G)
    r, c, v = po.csr_build(N, items, prefsum, op, og)
    B = bits.astype(np.int64)
    inter, ln, table = po.similarity(r, c, G)
    assert (inter.astype(np.int64) == B.T @ B).all() and (ln.astype(np.int64) == B.sum(0)).all()
    inter, ln, table = po.similarity(r, c, G, count_bp=True, node_lens=w)
    assert (inter.astype(np.int64) == B.T @ (B * w.astype(np.int64)[:, None])).all()
    i, j = 2, 5
    assert table[i, j] == np.float32(inter[i, j]) / np.float32(ln[i] + ln[j] - inter[i, j])


# ---- threshold parsing + TSV writers ----------------------------------------------------------------

def test_threshold_parse_and_strings():
    c, q = po.parse_thresholds("0,0.5,1.0", "1")
    assert [po.threshold_string(x) for x in q] == ["0", "0.5", "1"]
    assert [po.threshold_string(x) for x in c] == ["1", "1", "1"]
    with pytest.raises(ValueError):
        po.parse_thresholds("0,0.5", "1,2,3")
    with pytest.raises(ValueError):
        po.parse_thresholds("1.5", "1")


def test_tsv_shapes_from_survey_appendix_a():
    g, t, op, og, names = _tables("t_groups.gfa", "node")
    countable = po.abacus_by_total(t.n_items, t.items, t.id_prefsum, op, og, t.exclude)
    hist = po.construct_hist(countable, len(names))
    assert po.hist_table([("node", hist)]) == (
        "panacus\thist\ncount\tnode\n\t\n\t\n0\t5\n1\t0\n2\t10\n3\t0\n4\t0\n5\t0\n6\t0\n")
    cov, quo = po.parse_thresholds("0", "1")
    assert po.growth_table([("node", hist)], cov, quo) == (
        "panacus\tgrowth\ncount\tnode\ncoverage\t1\nquorum\t0\n0\tNaN\n1\t3\n2\t6\n3\t8\n4\t9\n5\t10\n6\t10\n")
    r, c, v = po.csr_build(t.n_items, t.items, t.id_prefsum, op, og, t.exclude)
    curve = po.calc_growth(r, c, len(names), cov[0], quo[0])
    assert po.ordered_growth_table("node", names, [curve], cov, quo) == (
        "panacus\tordered-growth\ncount\tnode\ncoverage\t1\nquorum\t0\n"
        "y#1\t2\ny#2\t5\ny#3\t8\ny#4\t9\ny#5\t10\nx\t10\n")


# ---- AbacusByGroup::to_tsv restatement (abacus.rs:1056-1178): no golden in the reference, so the oracle's writer is
# checked against an independent formulation -- the table recomputed straight from the GFA's P lines ----------------

@pytest.mark.parametrize("gfa,kw", [("t_groups.gfa", {}), ("chrM_test.gfa", {"groupby_sample": True}), ("cdbg.gfa", {})])
def test_table_writer_equals_direct_count(gfa, kw):
    path = os.path.join(GOLDEN, gfa)
    g = go.parse_gfa(path)
    mask = go.make_mask(g, **kw)
    op, og, names = go.path_order_arrays(mask, g)
    id2name = {v: k.decode() for k, v in g.node2id.items()}
    for count in ("node", "bp"):
        t = go.item_tables(g, mask, count)
        r, c, v = po.csr_build(t.n_items, t.items, t.id_prefsum, op, og, t.exclude)
        # occurrences of every node in every group, counted directly from the path steps
        occ = np.zeros((g.node_count + 1, len(names)), dtype=np.int64)
        for pid, grp in zip(op, og):
            for node, _ in g.path_steps[int(pid)]:
                occ[node, int(grp)] += 1
        lines = ["node\t" + "\t".join(names)]
        totals = ["node\ttotal"]
        for i in range(1, g.node_count + 1):
            w = g.node_lens[i] if count == "bp" else 1
            lines.append(id2name[i] + "".join(f"\t{int(x) * w}" for x in occ[i]))
            totals.append(f"{id2name[i]}\t{int((occ[i] > 0).sum())}")
        assert go.abacus_by_group_to_tsv(g, count, False, names, r, c, v, t.uncovered) == "\n".join(lines) + "\n"
        assert go.abacus_by_group_to_tsv(g, count, True, names, r, c, v, t.uncovered) == "\n".join(totals) + "\n"


def test_table_known_answer_t_groups():
    g = go.parse_gfa(os.path.join(GOLDEN, "t_groups.gfa"))
    mask = go.make_mask(g)
    op, og, names = go.path_order_arrays(mask, g)
    t = go.item_tables(g, mask, "bp")
    r, c, v = po.csr_build(t.n_items, t.items, t.id_prefsum, op, og, t.exclude)
    rows = go.abacus_by_group_to_tsv(g, "bp", False, names, r, c, v, t.uncovered).split("\n")
    assert rows[0] == "node\ty#1\ty#2\ty#3\ty#4\ty#5\tx"
    assert rows[1] == "1\t8\t0\t0\t0\t0\t8"        # segment 1 (8 bp) is on y#1 and x
    assert rows[9] == "9\t0\t0\t19\t0\t0\t19"      # segment 9 (19 bp) is on y#3 and x
    assert rows[2] == "2\t0\t0\t0\t0\t0\t0"        # segment 2 is on no path
    assert po.coverage_line_table([("node", [5, 0, 10, 0, 0, 0, 0])]).split("\n")[4:7] == ["1\t0", "2\t10", "3\t0"]


def test_ordered_growth_against_independent_witness():
    """The reference holds no golden vector for AbacusByGroup::calc_growth.  tests/golden/ordered_growth_witness.json is a
    brute-force evaluation of abacus.rs:1003-1010 straight from the P lines of the fixture GFAs
    (tests/golden/make_ordered_growth_witness.py: pure Python, no ItemTable, no CSR, nothing imported from oracle/); the
    oracle's chain GFA parser -> ItemTable -> CSR -> calc_growth must give the same curves for every grouping, count
    type and threshold pair."""
    import json
    from oracle import gfa_oracle as go
    d = json.load(open(os.path.join(GOLDEN, "ordered_growth_witness.json")))
    pairs = [tuple(p) for p in d["pairs"]]
    for case in d["cases"]:
        g = go.parse_gfa(os.path.join(GOLDEN, case["gfa"]))
        mask = go.make_mask(g, groupby_sample=case["grouping"] == "sample", groupby_haplotype=case["grouping"] == "haplotype")
        t = go.item_tables(g, mask, case["count"])
        op, og, names = go.path_order_arrays(mask, g)
        assert list(names) == case["groups"], (case["gfa"], case["grouping"])
        if t.n_items == 0:
            assert all(not any(cv) for cv in case["curves"])
            continue
        r, c, v = po.csr_build(t.n_items, t.items, t.id_prefsum, op, og, t.exclude)
        for (cov, q), want in zip(pairs, case["curves"]):
            got = po.calc_growth(r, c, len(names), po.absolute(cov), po.relative(q), count_bp=(case["count"] == "bp"),
                                 node_lens=g.node_lens)
            assert [int(x) for x in got] == want, (case["gfa"], case["grouping"], case["count"], cov, q)


def test_similarity_and_csr_against_independent_witness():
    """Like ordered growth, Similarity::set_table's integer loops and the AbacusByGroup CSR have no golden vector in the
    reference.  tests/golden/similarity_csr_witness.json counts both straight from the S / P lines of the fixture GFAs
    (tests/golden/make_similarity_csr_witness.py, nothing imported from oracle/); the oracle's chain GFA parser ->
    ItemTable -> cursor passes -> similarity loops must agree for every grouping and count type."""
    import json
    from oracle import gfa_oracle as go
    d = json.load(open(os.path.join(GOLDEN, "similarity_csr_witness.json")))
    for case in d["cases"]:
        g = go.parse_gfa(os.path.join(GOLDEN, case["gfa"]))
        mask = go.make_mask(g, groupby_sample=case["grouping"] == "sample", groupby_haplotype=case["grouping"] == "haplotype")
        t = go.item_tables(g, mask, case["count"])
        op, og, names = go.path_order_arrays(mask, g)
        assert list(names) == case["groups"], (case["gfa"], case["grouping"])
        G = len(names)
        if t.n_items == 0:
            assert not any(any(row) for row in case["inter"])
            continue
        r, c, v = po.csr_build(t.n_items, t.items, t.id_prefsum, op, og, t.exclude)
        if "r" in case:
            assert [int(x) for x in r] == case["r"] and [int(x) for x in c] == case["c"] and [int(x) for x in v] == case["v"], \
                (case["gfa"], case["grouping"], case["count"])
        inter, ln, _ = po.similarity(r, c, G, count_bp=(case["count"] == "bp"), node_lens=g.node_lens)
        assert inter.tolist() == case["inter"] and ln.tolist() == case["len"], (case["gfa"], case["grouping"], case["count"])
