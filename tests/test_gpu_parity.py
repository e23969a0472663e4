"""GPU parity tests: every hot-path entry point of libpanacus_b200.so (through the C ABI) against the
CPU oracle on the same inputs -- bit-exact (integer work).  Run with `-m gpu` on the B200 box."""
import json
import os

import numpy as np
import pytest

import panacus_b200 as pb
from panacus_b200 import synth
from oracle import gfa_oracle as go
from oracle import oracle as po

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
KATS = json.load(open(os.path.join(GOLDEN, "kats.json")))

PAIRS = [(1, 0.0), (2, 0.0), (1, 0.1), (2, 0.5), (4, 0.9), (1, 1.0), (3, 0.33)]


def oracle_all(bitmap, G, weights, pairs):
    """reference algorithm on the ItemTable derived from the bitmap -> dict of expected results"""
    N = bitmap.shape[0] - 1
    items, prefsum, op, og = po.bitmap_to_item_table(bitmap, G)
    countable = po.abacus_by_total(N, items, prefsum, op, og)
    r, c, v = po.csr_build(N, items, prefsum, op, og)
    exp = {
        "countable": countable,
        "hist": po.construct_hist(countable, G),
        "hist_bp": po.construct_hist_bps(countable, weights, G),
        "r": r, "c": c,
    }
    for (cov, q) in pairs:
        exp[("node", cov, q)] = po.calc_growth(r, c, G, po.absolute(cov), po.relative(q))
        exp[("bp", cov, q)] = po.calc_growth(r, c, G, po.absolute(cov), po.relative(q), count_bp=True, node_lens=weights)
    return exp


def cutoffs(G, pairs):
    cov = [max(1, c) for c, _ in pairs]
    thr = np.stack([pb.quorum_thresholds(G, q) for _, q in pairs])
    return cov, thr


def check_table(bitmap, G, weights, pairs=PAIRS):
    N = bitmap.shape[0] - 1
    exp = oracle_all(bitmap, G, weights, list(pairs) + [(1, 0.0), (2, 0.0), (2, 0.5)])
    with pb.DeviceAbacus(N, G) as a:
        a.upload(bitmap, weights)
        hc, hw, ct = a.hist(count=True, weight=True, countable=True)
        assert np.array_equal(ct, exp["countable"])
        assert np.array_equal(hc, exp["hist"])
        assert np.array_equal(hw, exp["hist_bp"])
        cov, thr = cutoffs(G, pairs)
        # general-quorum thresholds on both layouts: k_scan<quorum> (node-major) and k_gm_growth (group-major)
        for path in ("scan", "gm"):
            os.environ["PGX_QUORUM_PATH"] = path
            try:
                for weighted, key in ((False, "node"), (True, "bp")):
                    curves = a.ordered_growth(cov, thr, weighted=weighted)
                    for t, (c, q) in enumerate(pairs):
                        want = exp[(key, c, q)]
                        assert np.array_equal(curves[t].astype(np.float64), want), (path, key, c, q)
                h3, w3, cv3 = a.hist_ordered_growth(cov, thr, weighted=True, hist_count=True, hist_weight=True)
                assert np.array_equal(h3, hc) and np.array_equal(w3, hw) and np.array_equal(cv3, curves)
            finally:
                os.environ.pop("PGX_QUORUM_PATH", None)
        # fused hist + growth in one pass gives the same numbers
        hc2, hw2, cv2 = a.hist_ordered_growth(cov, thr, weighted=True, hist_count=True, hist_weight=True)
        assert np.array_equal(hc2, hc) and np.array_equal(hw2, hw)
        assert np.array_equal(cv2, a.ordered_growth(cov, thr, weighted=True))
        # q = 0 only (quorum_thr = NULL) uses the fast kernel alone
        c0 = a.ordered_growth([1, 2], None, weighted=False)
        assert np.array_equal(c0[0].astype(np.float64), exp[("node", 1, 0.0)])
        assert np.array_equal(c0[1].astype(np.float64), exp[("node", 2, 0.0)])
        # the reference-shaped convenience call
        got = a.calc_growth(pb.Threshold.absolute(2), pb.Threshold.relative(0.5), "bp")
        assert np.array_equal(got, exp[("bp", 2, 0.5)])
    return exp


# ---- reference fixtures ------------------------------------------------------------------------------

def fixture_bitmap(gfa, count, **mask_kw):
    g = go.parse_gfa(os.path.join(GOLDEN, gfa))
    mask = go.make_mask(g, **mask_kw)
    t = go.item_tables(g, mask, count)
    op, og, names = go.path_order_arrays(mask, g)
    G = len(names)
    bits = np.zeros((t.n_items + 1, G), dtype=np.uint8)
    for path_id, grp in zip(op, og):
        ids = t.items[int(t.id_prefsum[int(path_id)]):int(t.id_prefsum[int(path_id) + 1])].astype(np.int64)
        if t.exclude is not None:
            ids = ids[t.exclude[ids] == 0]
        bits[ids, int(grp)] = 1
    if count == "edge":
        weights = np.ones(t.n_items + 1, dtype=np.uint32)
    else:
        weights = np.array(g.node_lens, dtype=np.uint32)
    return g, t, op, og, names, bits, weights


@pytest.mark.parametrize("count", ["node", "bp", "edge"])
def test_chrM_against_reference_goldens(count):
    g, t, op, og, names, bits, weights = fixture_bitmap("chrM_test.gfa", count, groupby_sample=True)
    k = KATS["chrM_groupby_sample"]
    with pb.DeviceAbacus(t.n_items, len(names)) as a:
        a.upload(pb.pack_bits(bits), weights)
        hc, hw, ct = a.hist(count=True, weight=True, countable=True)
        assert list(ct[1:]) == k["countable_" + count] and ct[0] == 0xFFFFFFFF
        if count == "bp":
            assert list(hw) == k["bp"]      # abacus.rs:1630
        else:
            assert list(hc) == k[count]     # abacus.rs:1525 / 1579
        # ordered growth against the reference algorithm on the reference's own ItemTable
        r, c, v = po.csr_build(t.n_items, t.items, t.id_prefsum, op, og, t.exclude)
        cov, thr = cutoffs(len(names), PAIRS)
        curves = a.ordered_growth(cov, thr, weighted=(count == "bp"))
        for i, (cv, q) in enumerate(PAIRS):
            want = po.calc_growth(r, c, len(names), po.absolute(cv), po.relative(q), count_bp=(count == "bp"),
                                  node_lens=g.node_lens)
            assert np.array_equal(curves[i].astype(np.float64), want), (cv, q)
    if count == "node":
        assert list(curves[0]) == [89, 106, 140, 154]


def test_ordered_growth_against_independent_witness():
    """The CUDA path against tests/golden/ordered_growth_witness.json (brute force over the GFA's P lines, independent
    of the oracle): node-major and group-major kernels, every grouping / count type / threshold pair of the witness."""
    d = json.load(open(os.path.join(GOLDEN, "ordered_growth_witness.json")))
    pairs = [tuple(p) for p in d["pairs"]]
    for case in d["cases"]:
        g, t, op, og, names, bits, weights = fixture_bitmap(case["gfa"], case["count"], groupby_sample=case["grouping"] == "sample",
                                                            groupby_haplotype=case["grouping"] == "haplotype")
        assert list(names) == case["groups"]
        if t.n_items == 0:
            continue
        G = len(names)
        cov, thr = cutoffs(G, pairs)
        with pb.DeviceAbacus(t.n_items, G) as a:
            path_group = np.full(len(t.id_prefsum) - 1, -1, dtype=np.int64)
            path_group[op.astype(np.int64)] = og.astype(np.int64)
            a.build(t.items, t.id_prefsum, path_group, t.exclude)  # from the ItemTable, like the CLI
            a.upload(None, weights)
            for path in ("scan", "gm"):
                os.environ["PGX_QUORUM_PATH"] = path
                try:
                    cv = a.ordered_growth(cov, thr, weighted=(case["count"] == "bp"))
                finally:
                    os.environ.pop("PGX_QUORUM_PATH", None)
                assert [[int(x) for x in row] for row in cv] == case["curves"], (case["gfa"], case["grouping"], case["count"], path)
            ident = np.arange(G, dtype=np.uint32)[None, :]
            pg = a.permuted_growth(ident, cov, thr, weighted=(case["count"] == "bp"))
            assert [[int(x) for x in row] for row in pg[0]] == case["curves"]


@pytest.mark.parametrize("gfa", ["t_groups.gfa", "cdbg.gfa"])
def test_small_fixtures(gfa):
    for count in ("node", "bp", "edge"):
        g, t, op, og, names, bits, weights = fixture_bitmap(gfa, count)
        if t.n_items == 0:
            continue
        check_table(pb.pack_bits(bits), len(names), weights)


def test_chrM_subset_and_exclude():
    for kw in ({"subset": os.path.join(GOLDEN, "inclusion.bed1")},
               {"exclude": os.path.join(GOLDEN, "exclusion.bed3")},
               {"groupby_haplotype": True}):
        g, t, op, og, names, bits, weights = fixture_bitmap("chrM_test.gfa", "node", **kw)
        exp = check_table(pb.pack_bits(bits), len(names), weights, pairs=[(1, 0.0), (2, 0.5)])
        # the bitmap built from (items, exclude) reproduces the reference's own coverage vector
        want = po.abacus_by_total(t.n_items, t.items, t.id_prefsum, op, og, t.exclude)
        assert np.array_equal(exp["countable"], want)


# ---- random shapes: ragged widths, empty rows, full rows, every tile/tail combination -------------------

SHAPES = [(1, 1), (2, 3), (3, 64), (4, 65), (5, 63), (255, 100), (256, 128), (257, 129), (1000, 256),
          (1023, 300), (1030, 1024), (4099, 512), (515, 1100), (130, 2100), (70, 4500)]


@pytest.mark.parametrize("N,G", SHAPES)
def test_random_tables(N, G):
    rng = np.random.default_rng(N * 7919 + G)
    bits, bitmap, weights = synth.numpy_table(N, G, seed=N * 31 + G)
    # edge rows: empty, full, single first / last bit
    if N >= 4:
        bits[1] = 0
        bits[2] = 1
        bits[3] = 0
        bits[3, 0] = 1
        bits[4] = 0
        bits[4, G - 1] = 1
        bitmap = pb.pack_bits(bits)
    weights[1:] = rng.integers(0, 100001, N)
    if N >= 2:
        weights[2] = 0xFFFFFFFF  # forces the 32-bit carry path of the shared-memory accumulators
    check_table(bitmap, G, weights)


@pytest.mark.parametrize("N,G", [(300_000, 44), (150_000, 64), (120_000, 100), (100_000, 130), (90_000, 256), (40_000, 300), (777, 7)])
def test_lane_private_counter_kernel(N, G, monkeypatch):
    """k_scan_priv (lane-private counters instead of shared atomics, G <~ 300) against the oracle and against the
    atomics kernel: count and bp modes, several q = 0 thresholds with repeated coverage cutoffs (classes), per-item
    coverage output, and -- with the grid capped to 3 CTAs -- many tiles per CTA, so that the narrow counters wrap and
    carry many times and the ring wraps around."""
    bits, bitmap, weights = synth.numpy_table(N, G, seed=N + G)
    pairs = [(1, 0.0), (3, 0.0), (2, 0.0), (3, 0.0), (G, 0.0), (G + 1, 0.0)]
    exp = oracle_all(bitmap, G, weights, pairs)
    cov = [c for c, _ in pairs]
    with pb.DeviceAbacus(N, G) as a:
        a.upload(bitmap, weights)
        for grid in ("3", None):
            if grid:
                monkeypatch.setenv("PGX_SCAN_GRID", grid)
            else:
                monkeypatch.delenv("PGX_SCAN_GRID", raising=False)
            # "force": lane-private counters whenever they fit in one CTA; "off": the shared-atomics kernel; "auto": the
            # library's own choice (private counters only where two CTAs per SM fit)
            for mode in ("force", "off", "auto"):
                if mode == "auto":
                    monkeypatch.delenv("PGX_SCAN_PRIV", raising=False)
                else:
                    monkeypatch.setenv("PGX_SCAN_PRIV", "1" if mode == "force" else "2")
                hc, _, ct = a.hist(count=True, weight=False, countable=True)
                if mode != "auto":
                    assert ("k_scan_priv<u8>" in a.last_launch_info()) == (mode == "force"), a.last_launch_info()
                assert np.array_equal(hc, exp["hist"]) and np.array_equal(ct, exp["countable"])
                _, hw, _ = a.hist(count=False, weight=True)
                if mode != "auto":
                    assert ("k_scan_priv<u16>" in a.last_launch_info()) == (mode == "force"), a.last_launch_info()
                assert np.array_equal(hw, exp["hist_bp"])
                h2, _, cv = a.hist_ordered_growth(cov, None, weighted=False, hist_count=True, hist_weight=False)
                assert mode != "off" or "k_scan_priv" not in a.last_launch_info()
                assert not (mode == "force" and G <= 64) or "k_scan_priv<u8>" in a.last_launch_info(), a.last_launch_info()
                _, w2, cvw = a.hist_ordered_growth(cov, None, weighted=True, hist_count=False, hist_weight=True)
                assert np.array_equal(h2, exp["hist"]) and np.array_equal(w2, exp["hist_bp"])
                for t, (c, q) in enumerate(pairs):
                    assert np.array_equal(cv[t].astype(np.float64), exp[("node", c, q)]), (grid, mode, c)
                    assert np.array_equal(cvw[t].astype(np.float64), exp[("bp", c, q)]), (grid, mode, c)
                one = a.ordered_growth([2], None, weighted=False)  # growth only: no histogram bins at all
                assert np.array_equal(one[0].astype(np.float64), exp[("node", 2, 0.0)])
                if mode == "force" and G <= 256:
                    assert "k_scan_priv<u8>" in a.last_launch_info(), a.last_launch_info()
                if mode == "auto":
                    # hist-only private counters + shared atomics for the curves (default for bp sums where histogram +
                    # classes leave no room for two CTAs; PGX_SCAN_HYBRID=1 also when counting)
                    monkeypatch.setenv("PGX_SCAN_HYBRID", "1")
                    h4, _, cv4 = a.hist_ordered_growth(cov[:2], None, weighted=False, hist_count=True, hist_weight=False)
                    info = a.last_launch_info()
                    _, w4, cvw4 = a.hist_ordered_growth(cov[:2], None, weighted=True, hist_count=False, hist_weight=True)
                    monkeypatch.delenv("PGX_SCAN_HYBRID")
                    if G in (130, 256):
                        assert "hist-only" in info, info
                    assert np.array_equal(h4, exp["hist"]) and np.array_equal(w4, exp["hist_bp"])
                    for t, (c, q) in enumerate(pairs[:2]):
                        assert np.array_equal(cv4[t].astype(np.float64), exp[("node", c, q)]), (grid, c)
                        assert np.array_equal(cvw4[t].astype(np.float64), exp[("bp", c, q)]), (grid, c)


def test_dense_and_sparse_variants():
    for variant in ("dense", "sparse"):
        bits, bitmap, weights = synth.numpy_table(3000, 200, seed=5, variant=variant)
        check_table(bitmap, 200, weights)


def test_garbage_in_padding_bits_is_ignored():
    bits, bitmap, weights = synth.numpy_table(600, 70, seed=11)
    exp = oracle_all(bitmap, 70, weights, [(1, 0.0), (2, 0.5)])
    dirty = bitmap.copy()
    dirty[:, 1] |= np.uint64(0xFFFFFFFFFFFFFFC0)  # bits 70..127
    dirty[0, :] = np.uint64(0xFFFFFFFFFFFFFFFF)   # dummy row
    with pb.DeviceAbacus(600, 70) as a:
        a.upload(dirty, weights)
        hc, hw, ct = a.hist(True, True, True)
        assert np.array_equal(hc, exp["hist"]) and np.array_equal(hw, exp["hist_bp"])
        cov, thr = cutoffs(70, [(1, 0.0), (2, 0.5)])
        cv = a.ordered_growth(cov, thr)
        assert np.array_equal(cv[0].astype(np.float64), exp[("node", 1, 0.0)])
        assert np.array_equal(cv[1].astype(np.float64), exp[("node", 2, 0.5)])


def test_many_thresholds_chunking():
    bits, bitmap, weights = synth.numpy_table(900, 96, seed=3)
    pairs = [(1 + (i % 5), [0.0, 0.25, 0.5, 0.75, 1.0][i % 5] if i % 2 else 0.0) for i in range(19)]
    exp = oracle_all(bitmap, 96, weights, pairs)
    with pb.DeviceAbacus(900, 96) as a:
        a.upload(bitmap, weights)
        cov, thr = cutoffs(96, pairs)
        cv = a.ordered_growth(cov, thr, weighted=True)
        for t, (c, q) in enumerate(pairs):
            assert np.array_equal(cv[t].astype(np.float64), exp[("bp", c, q)]), (t, c, q)


# ---- permuted growth (group-major kernels) -----------------------------------------------------------------

@pytest.mark.parametrize("kernel", ["table", "old"])
@pytest.mark.parametrize("N,G", [(5, 3), (300, 64), (1000, 100), (2049, 257), (700, 1030)])
def test_permuted_growth(N, G, kernel, monkeypatch):
    """general thresholds on k_gm_quorum (mask table in shared memory; default) and on k_gm_growth<P,true> (PGX_GM_QUORUM=old)"""
    if kernel == "old":
        monkeypatch.setenv("PGX_GM_QUORUM", "old")
    bits, bitmap, weights = synth.numpy_table(N, G, seed=N + G)
    orders = synth.random_orders(5, G, seed=99)
    orders[0] = np.arange(G)  # identity: must equal the node-major kernels
    pairs = [(1, 0.0), (2, 0.5), (4, 0.9)]
    cov, thr = cutoffs(G, pairs)
    with pb.DeviceAbacus(N, G) as a:
        a.upload(bitmap, weights)
        for weighted in (False, True):
            got = a.permuted_growth(orders, cov, thr, weighted=weighted)
            if kernel == "table" and G > 1:
                assert "k_gm_quorum" in a.last_launch_info()
            assert np.array_equal(got[0], a.ordered_growth(cov, thr, weighted=weighted))
            for p in range(orders.shape[0]):
                pbits = bits[:, orders[p]]  # the abacus the reference rebuilds under --order
                exp = oracle_all(pb.pack_bits(pbits), G, weights, pairs)
                for t, (c, q) in enumerate(pairs):
                    want = exp[("bp" if weighted else "node", c, q)]
                    assert np.array_equal(got[p, t].astype(np.float64), want), (p, c, q, weighted)
            one = a.ordered_growth(cov, thr, col_order=orders[3], weighted=weighted)
            assert np.array_equal(one, got[3])


def test_permuted_growth_many_mixed_thresholds():
    """3 q = 0 thresholds (one rides along with k_gm_quorum, two go to the HBM-bound kernel) + 7 general ones (two launches
    of <= 4), interleaved, with coverage cutoffs; caller-supplied cutoffs beyond G + 1 never count"""
    N, G = 1500, 90
    bits, bitmap, weights = synth.numpy_table(N, G, seed=77)
    pairs = [(1, 0.0), (2, 0.25), (1, 1.0), (2, 0.0), (3, 0.5), (1, 0.05), (5, 0.9), (3, 0.0), (2, 0.75), (1, 0.6)]
    cov, thr = cutoffs(G, pairs)
    orders = synth.random_orders(3, G, seed=5)
    with pb.DeviceAbacus(N, G) as a:
        a.upload(bitmap, weights)
        for weighted in (False, True):
            got = a.permuted_growth(orders, cov, thr, weighted=weighted)
            for p in range(orders.shape[0]):
                exp = oracle_all(pb.pack_bits(bits[:, orders[p]]), G, weights, pairs)
                for t, (c, q) in enumerate(pairs):
                    assert np.array_equal(got[p, t].astype(np.float64), exp[("bp" if weighted else "node", c, q)]), (p, c, q)
        never = np.full((1, G), 2 ** 31, dtype=np.uint32)
        assert not a.permuted_growth(orders[:1], [1], never).any()


@pytest.mark.parametrize("kernel", ["union", "old"])
@pytest.mark.parametrize("N,G", [(1500, 90), (4000, 300), (70, 33), (16500, 40)])
def test_permuted_growth_q0_only(N, G, kernel, monkeypatch):
    """q = 0 thresholds only (the permutation-sampled union / coverage >= c growth): k_gm_union (two columns per thread;
    default) and the first-generation k_gm_growth<.,false> (PGX_GM_QUORUM=old), with 1, 2 and 4 + 1 thresholds differing
    in their coverage cutoff; bp sums on the weight-sorted copy, with weights that leave both single-weight and mixed
    64-item columns; odd and even numbers of columns, more than one CTA per order"""
    if kernel == "old":
        monkeypatch.setenv("PGX_GM_QUORUM", "old")
    bits, bitmap, weights = synth.numpy_table(N, G, seed=N)
    rng = np.random.default_rng(N)
    weights = np.where(rng.random(N + 1) < 0.6, 1, rng.integers(1, 2 ** 32, N + 1, dtype=np.uint64)).astype(np.uint32)
    weights[0] = 0
    orders = synth.random_orders(3, G, seed=11)
    with pb.DeviceAbacus(N, G) as a:
        a.upload(bitmap, weights)
        for covs in ([1], [2, 1], [1, 2, 3, 4, 5]):
            pairs = [(c, 0.0) for c in covs]
            for weighted in (False, True):
                got = a.permuted_growth(orders, covs, None, weighted=weighted)
                assert ("k_gm_union" if kernel == "union" else "k_gm_growth") in a.last_launch_info()
                for p in range(orders.shape[0]):
                    exp = oracle_all(pb.pack_bits(bits[:, orders[p]]), G, weights, pairs)
                    for t, (c, q) in enumerate(pairs):
                        assert np.array_equal(got[p, t].astype(np.float64), exp[("bp" if weighted else "node", c, q)]), (p, c, weighted)


# ---- similarity -------------------------------------------------------------------------------------------

@pytest.mark.parametrize("N,G", [(3000, 256), (70_000, 300), (100_001, 1024), (5000, 700), (127, 256), (449, 512), (900, 70)])
def test_similarity_tensor_core_kernel(N, G, monkeypatch):
    """k_sim_mma (tcgen05.mma.kind::i8 on the bits expanded to u8, s32 accumulators in TMEM) against the oracle's
    restatement of Similarity::set_table and against the AND / POPC kernel: full square (upper tiles + mirror), row
    blocks, upper-triangle blocks; ragged shapes (groups and items that do not fill a tile / a stage)."""
    bits, bitmap, weights = synth.numpy_table(N, G, seed=N * 3 + G)
    items, prefsum, op, og = po.bitmap_to_item_table(bitmap, G)
    r, c, _ = po.csr_build(N, items, prefsum, op, og)
    want, want_len, _ = po.similarity(r, c, G)
    with pb.DeviceAbacus(N, G) as a:
        a.upload(bitmap, weights)
        monkeypatch.setenv("PGX_SIM", "mma")
        inter, ln = a.similarity()
        assert "k_sim_mma" in a.last_launch_info()
        assert np.array_equal(inter, want) and np.array_equal(ln, want_len)
        lo, hi = min(64, G // 3), min(G, 300)
        part, ln2 = a.similarity(row_begin=lo, row_end=hi)
        assert np.array_equal(part, want[lo:hi]) and np.array_equal(ln2, want_len)
        up, _ = a.similarity(row_begin=lo, row_end=hi, upper=True)
        for x in range(lo, hi):  # columns >= the row (what a sharded run reads) are computed, columns < row_begin are zero
            assert np.array_equal(up[x - lo, x:], want[x, x:]) and not up[x - lo, :lo].any()
        monkeypatch.setenv("PGX_SIM", "csa")
        inter2, _ = a.similarity()
        assert np.array_equal(inter2, want)


@pytest.mark.parametrize("variant", ["csa", "plain"])
@pytest.mark.parametrize("N,G", [(10, 2), (500, 64), (3000, 70), (1500, 130), (70000, 40)])
def test_similarity(N, G, variant, monkeypatch):
    """unweighted intersections on the carry-save kernel (default) and on the plain AND + POPC kernel (PGX_SIM=plain)"""
    if variant == "plain":
        monkeypatch.setenv("PGX_SIM", "plain")
    bits, bitmap, weights = synth.numpy_table(N, G, seed=2 * N + G)
    bits[1, :] = 1
    bitmap = pb.pack_bits(bits)
    exp = oracle_all(bitmap, G, weights, [])
    with pb.DeviceAbacus(N, G) as a:
        a.upload(bitmap, weights)
        for weighted in (False, True):
            inter_o, len_o, table_o = po.similarity(exp["r"], exp["c"], G, count_bp=weighted, node_lens=weights)
            inter, ln = a.similarity(weighted=weighted)
            assert np.array_equal(inter, inter_o) and np.array_equal(ln, len_o)
            if not weighted:
                assert a.last_launch_info() == f"k_gm_similarity<{variant}>"
            # f32 Jaccard exactly as similarity.rs:153-163
            table = inter.astype(np.float32) / (ln[:, None] + ln[None, :] - inter).astype(np.float32)
            assert np.array_equal(table, table_o)
            # upper-triangle sharding: a block from its diagonal rightwards, the columns left of it stay zero
            for lo_u, hi_u in ((0, G), (G // 3, max(G // 3 + 1, 2 * G // 3)), (G - 1, G), (G // 2, G // 2)):
                up, ln_u = a.similarity(weighted=weighted, row_begin=lo_u, row_end=hi_u, upper=True)
                want_u = inter_o[lo_u:hi_u].copy()
                want_u[:, :lo_u] = 0
                assert np.array_equal(up, want_u) and np.array_equal(ln_u, len_o), (lo_u, hi_u)
            # row-block sharding (what each GPU computes in the 8-GPU configuration)
            lo, hi = G // 3, max(G // 3 + 1, 2 * G // 3)
            part, ln2 = a.similarity(weighted=weighted, row_begin=lo, row_end=hi)
            assert np.array_equal(part, inter_o[lo:hi]) and np.array_equal(ln2, len_o)


# ---- device-side build from ItemTables -------------------------------------------------------------------

def test_scatter_build_matches_reference_incidence():
    g, t, op, og, names, bits, weights = fixture_bitmap("chrM_test.gfa", "node", groupby_sample=True,
                                                        exclude=os.path.join(GOLDEN, "exclusion.bed3"))
    with pb.DeviceAbacus(t.n_items, len(names)) as a:
        for path_id, grp in zip(op, og):
            sl = t.items[int(t.id_prefsum[int(path_id)]):int(t.id_prefsum[int(path_id) + 1])]
            a.scatter(sl, int(grp), t.exclude)
        assert np.array_equal(a.download(), pb.pack_bits(bits))
        hc, _, ct = a.hist(True, False, True)
        want = po.abacus_by_total(t.n_items, t.items, t.id_prefsum, op, og, t.exclude)
        assert np.array_equal(ct, want)
        with pytest.raises(pb.PgxError):
            a.scatter(np.array([t.n_items + 1], dtype=np.uint64), 0)


def test_whole_table_build_matches_reference_incidence():
    for kw in ({"groupby_sample": True}, {"groupby_haplotype": True, "exclude": os.path.join(GOLDEN, "exclusion.bed3")},
               {"subset": os.path.join(GOLDEN, "inclusion.bed3")}):
        for count in ("node", "edge"):
            g, t, op, og, names, bits, weights = fixture_bitmap("chrM_test.gfa", count, **kw)
            path_group = np.full(len(t.id_prefsum) - 1, -1, dtype=np.int64)  # paths outside the order are not counted
            path_group[op.astype(np.int64)] = og.astype(np.int64)
            with pb.DeviceAbacus(t.n_items, len(names)) as a:
                a.build(t.items, t.id_prefsum, path_group, t.exclude)
                assert np.array_equal(a.download(), pb.pack_bits(bits)), (kw, count)
                _, _, ct = a.hist(True, False, True)
                assert np.array_equal(ct, po.abacus_by_total(t.n_items, t.items, t.id_prefsum, op, og, t.exclude))
    # random table with empty paths and repeated visits; one path left out
    rng = np.random.default_rng(5)
    N, P, G = 5000, 40, 13
    lens = rng.integers(0, 800, P)
    lens[[3, 4, 17]] = 0
    items = rng.integers(1, N + 1, int(lens.sum())).astype(np.uint64)
    prefsum = np.concatenate([[0], np.cumsum(lens)]).astype(np.uint64)
    path_group = rng.integers(0, G, P).astype(np.int64)
    path_group[7] = -1
    want = np.zeros((N + 1, G), dtype=np.uint8)
    for p in range(P):
        if path_group[p] >= 0:
            want[items[int(prefsum[p]):int(prefsum[p + 1])].astype(np.int64), path_group[p]] = 1
    with pb.DeviceAbacus(N, G) as a:
        a.build(items, prefsum, path_group)
        assert np.array_equal(a.download(), pb.pack_bits(want))
        with pytest.raises(pb.PgxError):
            a.build(np.array([N + 1], dtype=np.uint64), np.array([0, 1], dtype=np.uint64), np.array([0], dtype=np.int64))
        with pytest.raises(pb.PgxError):
            a.build(np.array([1], dtype=np.uint64), np.array([0, 1], dtype=np.uint64), np.array([G], dtype=np.int64))
        with pytest.raises(pb.PgxError):
            a.build(np.array([1, 2], dtype=np.uint64), np.array([0, 1], dtype=np.uint64), np.array([0], dtype=np.int64))


def test_build_u32_pinned_and_chunked_pipeline():
    """pgx_abacus_build / pgx_abacus_build_u32 from pageable and page-locked tables, over more than one 16 Mi-step chunk
    (double-buffered upload overlapping k_build): all four give the bitmap the reference's per-path loops define."""
    rng = np.random.default_rng(11)
    N, P, G = 300_000, 70, 70
    lens = rng.integers(100_000, 400_000, P)
    lens[[5, 6, 40]] = 0
    lens[9] = 17_000_000  # one path longer than a chunk: its steps straddle the staging buffers
    items = rng.integers(1, N + 1, int(lens.sum())).astype(np.uint64)
    prefsum = np.concatenate([[0], np.cumsum(lens)]).astype(np.uint64)
    path_group = rng.permutation(P).astype(np.int64) % G
    path_group[3] = -1
    exclude = (rng.random(N + 1) < 0.05).astype(np.uint8)
    want = np.zeros((N + 1, G), dtype=np.uint8)
    for p in range(P):
        if path_group[p] >= 0:
            ids = items[int(prefsum[p]):int(prefsum[p + 1])].astype(np.int64)
            want[ids[exclude[ids] == 0], path_group[p]] = 1
    want = pb.pack_bits(want)
    items32 = items.astype(np.uint32)
    pin64 = pb.pinned_empty(items.size, np.uint64)
    pin64[:] = items
    pin32 = pb.pinned_empty(items.size, np.uint32)
    pin32[:] = items32
    with pb.DeviceAbacus(N, G) as a:
        for src in (items, items32, pin64, pin32):
            a.clear()
            a.build(src, prefsum, path_group, exclude)
            assert np.array_equal(a.download(), want), (src.dtype, a.last_launch_info())
        assert "pinned" in a.last_launch_info() and "u32" in a.last_launch_info()
        a.build(items, prefsum, path_group, exclude)  # ORs into the existing bits: idempotent
        assert np.array_equal(a.download(), want)
        bad = items32.copy()
        bad[-1] = N + 1
        with pytest.raises(pb.PgxError):
            a.build(bad, prefsum, path_group)


def test_copy_rows_and_results_into_pinned_memory():
    N, G = 50_000, 130
    bits, bitmap, weight = synth.numpy_table(N, G, seed=21)
    cov, thr = cutoffs(G, [(1, 0.0), (2, 0.5)])
    with pb.DeviceAbacus(N, G) as full:
        full.upload(bitmap, weight)
        lo, n = 12_345, 20_000
        with pb.DeviceAbacus(n, G) as part:
            part.copy_rows_from(full, lo)
            got = part.download()
            assert np.array_equal(got[1:], bitmap[lo:lo + n, :got.shape[1]]) and not got[0].any()
            _, hw, _ = part.hist(count=False, weight=True)
            assert int(hw.sum()) == int(weight[lo:lo + n].sum())
            with pytest.raises(pb.PgxError):
                part.copy_rows_from(full, N - n + 2)
        orders = synth.random_orders(3, G, seed=4)
        want = full.permuted_growth(orders, cov, thr)
        out = pb.pinned_empty(want.shape, np.uint64)
        assert full.permuted_growth(orders, cov, thr, out=out) is out and np.array_equal(out, want)
        inter, ln = full.similarity()
        outi = pb.pinned_empty((G, G), np.uint64)
        i2, l2 = full.similarity(out_inter=outi)
        assert i2 is outi and np.array_equal(outi, inter) and np.array_equal(l2, ln)
        # kernel timing: events around the hot kernels of a call
        full.set_timing(True)
        full.kernel_time_ms()
        full.similarity()
        ms, sections = full.kernel_time_ms()
        assert sections == 1 and ms > 0
        full.hist_ordered_growth(cov, None)
        ms, sections = full.kernel_time_ms()
        assert sections >= 1 and ms > 0
        full.set_timing(False)
        full.hist()
        assert full.kernel_time_ms() == (0.0, 0)


def test_config2_full_against_oracle():
    """BASELINE.json configs[1] in full: 1M items x 256 groups, count = node, hist + ordered growth for the three
    threshold pairs, every value against the C oracle (ItemTable -> coverage -> CSR -> calc_growth)."""
    import torch
    N, G = 1_000_000, 256
    bitmap, weight = synth.torch_table(N, G, seed=synth.SEED_BASE + 2)
    host = bitmap.cpu().numpy().view(np.uint64)
    wh = weight.cpu().numpy().view(np.uint32)
    pairs = [(1, 0.0), (2, 0.5), (4, 0.9)]
    exp = oracle_all(host, G, wh, pairs)
    cov, thr = cutoffs(G, pairs)
    with pb.DeviceAbacus(N, G) as a:
        a.adopt_device(bitmap.data_ptr(), weight.data_ptr(), keepalive=(bitmap, weight))
        hc, hw, cv = a.hist_ordered_growth(cov, thr, weighted=False, hist_count=True, hist_weight=True)
        assert np.array_equal(hc, exp["hist"]) and np.array_equal(hw, exp["hist_bp"])
        cvw = a.ordered_growth(cov, thr, weighted=True)
        for t, (c, q) in enumerate(pairs):
            assert np.array_equal(cv[t].astype(np.float64), exp[("node", c, q)]), (c, q)
            assert np.array_equal(cvw[t].astype(np.float64), exp[("bp", c, q)]), (c, q)
    del torch


def test_config3_shape_random_orders_against_oracle():
    """BASELINE.json configs[2]'s shape (512 groups, pairs (1,0) (2,0.5) (4,0.9)) on a 200k-item slice of the bench table:
    growth under 3 random group orders against the C oracle run on the abacus rebuilt under each order (what the
    reference does for `--order`, abacus.rs:324-326)."""
    N, G = 200_000, 512
    bitmap, weight = synth.torch_table(N, G, seed=synth.SEED_BASE + 3)
    host = bitmap.cpu().numpy().view(np.uint64)
    wh = weight.cpu().numpy().view(np.uint32)
    pairs = [(1, 0.0), (2, 0.5), (4, 0.9)]
    cov, thr = cutoffs(G, pairs)
    orders = synth.random_orders(3, G, seed=synth.SEED_BASE + 3)
    bits = np.unpackbits(host.view(np.uint8).reshape(N + 1, -1), axis=1, bitorder="little")[:, :G]
    with pb.DeviceAbacus(N, G) as a:
        a.adopt_device(bitmap.data_ptr(), weight.data_ptr(), keepalive=(bitmap, weight))
        pg = a.permuted_growth(orders, cov, thr, weighted=False)
        pgw = a.permuted_growth(orders[:1], cov, thr, weighted=True)
    for p in range(3):
        items, prefsum, op, og = po.bitmap_to_item_table(pb.pack_bits(bits[:, orders[p]]), G)
        r, c, _ = po.csr_build(N, items, prefsum, op, og)
        for t, (cv, q) in enumerate(pairs):
            assert np.array_equal(pg[p, t].astype(np.float64), po.calc_growth(r, c, G, po.absolute(cv), po.relative(q))), (p, cv, q)
            if p == 0:
                want = po.calc_growth(r, c, G, po.absolute(cv), po.relative(q), count_bp=True, node_lens=wh)
                assert np.array_equal(pgw[0, t].astype(np.float64), want), ("bp", cv, q)


def test_csr_build_matches_reference_r_c_v():
    """AbacusByGroup {r, c, v}: device CSR (bitmap popcounts + scan, bit positions, one atomic per step) against the
    reference's two cursor passes (abacus.rs:859-986) restated in oracle/."""
    for kw in ({}, {"groupby_sample": True}, {"groupby_haplotype": True, "exclude": os.path.join(GOLDEN, "exclusion.bed3")},
               {"subset": os.path.join(GOLDEN, "inclusion.bed3")}):
        for count in ("node", "edge"):
            g, t, op, og, names, bits, weights = fixture_bitmap("chrM_test.gfa", count, **kw)
            path_group = np.full(len(t.id_prefsum) - 1, -1, dtype=np.int64)
            path_group[op.astype(np.int64)] = og.astype(np.int64)
            r0, c0, v0 = po.csr_build(t.n_items, t.items, t.id_prefsum, op, og, t.exclude)
            with pb.DeviceAbacus(t.n_items, len(names)) as a:
                a.build(t.items, t.id_prefsum, path_group, t.exclude)
                r, c, v = a.csr(t.items, t.id_prefsum, path_group, t.exclude)
                assert np.array_equal(r, r0) and np.array_equal(c, c0) and np.array_equal(v, v0), (kw, count)
                r2, c2, v2 = a.csr(values=False)
                assert np.array_equal(r2, r0) and np.array_equal(c2, c0) and v2 is None
    # random table: many paths per group, G > 64 (multi-word ranks), repeated visits, empty paths, one path left out
    rng = np.random.default_rng(11)
    for N, P, G in ((3000, 60, 13), (2000, 300, 150), (1, 3, 2)):
        lens = rng.integers(0, 400, P)
        lens[rng.integers(0, P, 3)] = 0
        items = rng.integers(1, N + 1, int(lens.sum())).astype(np.uint64)
        prefsum = np.concatenate([[0], np.cumsum(lens)]).astype(np.uint64)
        path_group = np.sort(rng.integers(0, G, P)).astype(np.int64)  # a group's paths are contiguous (abacus.rs:310-347)
        path_group[P // 2] = -1
        exclude = (rng.random(N + 1) < 0.05).astype(np.uint8)
        keep = np.flatnonzero(path_group >= 0)
        op, og = keep.astype(np.uint64), path_group[keep].astype(np.uint64)
        for ex in (None, exclude):
            r0, c0, v0 = po.csr_build(N, items, prefsum, op, og, ex)
            with pb.DeviceAbacus(N, G) as a:
                a.build(items, prefsum, path_group, ex)
                r, c, v = a.csr(items, prefsum, path_group, ex)
                assert np.array_equal(r, r0) and np.array_equal(c, c0) and np.array_equal(v, v0), (N, P, G, ex is None)
                assert int(v.sum()) == sum(int(lens[p]) for p in keep) - (0 if ex is None else
                                                                           int(sum(ex[items[int(prefsum[p]):int(prefsum[p + 1])].astype(np.int64)].sum() for p in keep)))
    # a table the bitmap was not built from is refused
    with pb.DeviceAbacus(10, 4) as a:
        a.build(np.array([1, 2], dtype=np.uint64), np.array([0, 2], dtype=np.uint64), np.array([0], dtype=np.int64))
        with pytest.raises(pb.PgxError):
            a.csr(np.array([3], dtype=np.uint64), np.array([0, 1], dtype=np.uint64), np.array([1], dtype=np.int64))


def test_argument_errors():
    with pytest.raises(pb.PgxError):
        pb.DeviceAbacus(10, 0)
    with pb.DeviceAbacus(10, 8) as a:
        with pytest.raises(pb.PgxError):
            a.ordered_growth([0], None)            # coverage cutoff must already be clamped to >= 1
        with pytest.raises(pb.PgxError):
            a.ordered_growth([1], None, col_order=np.zeros(8, dtype=np.uint32))  # not a permutation
        with pytest.raises(pb.PgxError):
            a.similarity(row_begin=5, row_end=3)
        hc, _, _ = a.hist()                        # empty abacus: every item has coverage 0
        assert hc[0] == 10 and hc[1:].sum() == 0


# ---- full-size properties (BASELINE.json shapes; no CPU oracle at this size) ---------------------------------

def test_full_size_properties():
    import torch
    N, G = 2_000_000, 1024
    bitmap, weight = synth.torch_table(N, G, seed=synth.SEED_BASE + 1)
    torch.cuda.synchronize()
    with pb.DeviceAbacus(N, G) as a:
        a.adopt_device(bitmap.data_ptr(), weight.data_ptr(), keepalive=(bitmap, weight))
        pairs = [(1, 0.0), (2, 0.5), (4, 0.9)]
        cov, thr = cutoffs(G, pairs)
        hc, hw, cv = a.hist_ordered_growth(cov, thr, weighted=False, hist_count=True, hist_weight=True)
        _, _, cvw = a.hist_ordered_growth(cov, thr, weighted=True, hist_count=False, hist_weight=False)
        wsum = int(weight.to(torch.int64).sum().item())
        assert int(hc.sum()) == N and int(hw.sum()) == wsum
        # every item with >= 1 group is counted once all groups are in: curve end = N - hist[0]
        assert int(cv[0, -1]) == N - int(hc[0]) and int(cvw[0, -1]) == wsum - int(hw[0])
        assert (np.diff(cv[0].astype(np.int64)) >= 0).all()
        # c = 2, q = 0.5 at the last column: all items with cov >= 2 whose final verdict holds; bounded by cov >= 2
        assert int(cv[1, -1]) <= int(hc[2:].sum()) and int(cv[2, -1]) <= int(hc[4:].sum())
        # popcount cross-check of the histogram against torch on the same device buffer
        sub = bitmap[: 200_001]
        bits = ((sub.unsqueeze(-1) >> torch.arange(64, device="cuda")) & 1).sum(dim=(1, 2))
        ref_hist = torch.bincount(bits[1:], minlength=G + 1).cpu().numpy()
        with pb.DeviceAbacus(200_000, G) as b:
            b.adopt_device(sub.data_ptr(), None, keepalive=sub)
            hb, _, _ = b.hist()
            assert np.array_equal(hb.astype(np.int64), ref_hist)
        # the group-major kernels (independent implementation) agree with the node-major pass
        ident = np.arange(G, dtype=np.uint32)[None, :]
        gm = a.permuted_growth(ident, cov, thr, weighted=True)
        assert np.array_equal(gm[0], cvw)
        # a sub-sample small enough for the CPU oracle, taken from the same device table
        M = 20_000
        host = bitmap[: M + 1].cpu().numpy().view(np.uint64)
        wh = weight[: M + 1].cpu().numpy().view(np.uint32)
        exp = oracle_all(host, G, wh, pairs)
        with pb.DeviceAbacus(M, G) as b:
            b.upload(host, wh)
            cvs = b.ordered_growth(cov, thr, weighted=True)
            for t, (c, q) in enumerate(pairs):
                assert np.array_equal(cvs[t].astype(np.float64), exp[("bp", c, q)])


def test_large_permuted_growth_and_similarity_properties():
    """BASELINE.json configs 3 / 4 at (reduced but HBM-sized) shapes: size-independent invariants that tie the
    group-major kernels to the pinned histogram path."""
    import torch
    N, G = 1_000_000, 512
    bitmap, weight = synth.torch_table(N, G, seed=synth.SEED_BASE + 3)
    torch.cuda.synchronize()
    with pb.DeviceAbacus(N, G) as a:
        a.adopt_device(bitmap.data_ptr(), weight.data_ptr(), keepalive=(bitmap, weight))
        hc, hw, ct = a.hist(True, True, True)
        pairs = [(1, 0.0), (2, 0.5), (4, 0.9)]
        cov, thr = cutoffs(G, pairs)
        orders = synth.random_orders(6, G, seed=synth.SEED_BASE + 3)
        orders[0] = np.arange(G)
        for weighted, h in ((False, hc), (True, hw)):
            pg = a.permuted_growth(orders, cov, thr, weighted=weighted)
            ref = a.ordered_growth(cov, thr, weighted=weighted)
            assert np.array_equal(pg[0], ref)                      # identity order == group order
            total = int(h[1:].sum())
            for p in range(orders.shape[0]):
                assert int(pg[p, 0, -1]) == total                   # union of all groups is order independent
                assert (np.diff(pg[p, 0].astype(np.int64)) >= 0).all()
                # last column of (c, q): items with coverage >= c whose final verdict holds; quorum only removes
                assert int(pg[p, 1, -1]) <= int(h[2:].sum()) and int(pg[p, 2, -1]) <= int(h[4:].sum())
            # q = 1.0: an item keeps counting exactly while its groups form a prefix of the order (the reference
            # tests the quorum against the last group that contained the item, abacus.rs:1007-1010), so the curve
            # is non-increasing after position 0 and never below the core (items present in every group)
            thr1 = np.stack([pb.quorum_thresholds(G, 1.0)])
            core = a.permuted_growth(orders[:3], [1], thr1, weighted=weighted)
            for p in range(3):
                assert (np.diff(core[p, 0].astype(np.int64)) <= 0).all() and int(core[p, 0, -1]) >= int(h[G])
        # first point of every curve = size of the first group of that order = len[] of the similarity path
        inter, ln = a.similarity(weighted=False)
        pg = a.permuted_growth(orders, [1], None, weighted=False)
        assert [int(pg[p, 0, 0]) for p in range(6)] == [int(ln[orders[p, 0]]) for p in range(6)]
        # similarity invariants: symmetric, diagonal = len, bounded by min(len), total incidences match the hist
        assert np.array_equal(inter, inter.T) and np.array_equal(np.diag(inter), ln)
        assert (inter <= np.minimum(ln[:, None], ln[None, :])).all()
        assert int(ln.sum()) == int((np.arange(G + 1, dtype=np.uint64) * hc).sum())
        # sum over pairs = sum over items of coverage^2 (every item contributes cov x cov ordered pairs)
        cov2 = np.bincount(ct[1:].astype(np.int64), minlength=G + 1).astype(object)
        assert int(inter.astype(object).sum()) == int(sum(int(c) * int(c) * int(cov2[c]) for c in range(G + 1)))
        interw, lnw = a.similarity(weighted=True)
        wnp = weight.cpu().numpy().view(np.uint32).astype(np.int64)
        ctn = ct.astype(np.int64)
        assert int(lnw.astype(object).sum()) == int((wnp[1:] * ctn[1:]).sum())
        assert int(interw.astype(object).sum()) == int((wnp[1:] * ctn[1:] * ctn[1:]).sum())
        part, _ = a.similarity(weighted=True, row_begin=100, row_end=229)
        assert np.array_equal(part, interw[100:229])


@pytest.mark.parametrize("N,G", [(0, 5), (1, 1), (3, 17000), (40, 9000), (100, 65)])
def test_extreme_shapes(N, G):
    """no items at all, a single cell, very wide rows (row > one pipeline stage of 256 items), many thresholds"""
    rng = np.random.default_rng(N + G)
    bits = (rng.random((N + 1, G)) < 0.3).astype(np.uint8)
    bits[0] = 0
    weights = rng.integers(0, 1000, N + 1).astype(np.uint32)
    pairs = [(1, 0.0), (2, 0.4), (1, 1.0)] if G > 1000 else [(1 + i % 3, [0.0, 0.2, 0.6, 1.0][i % 4]) for i in range(11)]
    if N == 0:
        with pb.DeviceAbacus(0, G) as a:
            a.upload(pb.pack_bits(bits), weights)
            hc, hw, ct = a.hist(True, True, True)
            assert not hc.any() and not hw.any() and ct[0] == 0xFFFFFFFF
            cov, thr = cutoffs(G, pairs)
            assert not a.ordered_growth(cov, thr, weighted=True).any()
            assert not a.permuted_growth(np.arange(G, dtype=np.uint32)[None, ::-1].copy(), cov, thr).any()
            inter, ln = a.similarity()
            assert not inter.any() and not ln.any()
        return
    check_table(pb.pack_bits(bits), G, weights, pairs=pairs)


# ---- bit-sliced vertical counters (k_scan_vert: G <= 64, counting) -------------------------------------------------

@pytest.mark.parametrize("N,G", [(3, 5), (4095, 1), (4096, 7), (4097, 31), (12_289, 32), (50_000, 33), (300_000, 44),
                                 (100_001, 63), (200_003, 64), (2, 65), (2047, 70), (2049, 100), (150_001, 127), (90_000, 128)])
def test_vertical_counter_kernel(N, G, monkeypatch):
    """k_scan_vert (carry-save adders over one-hot words instead of atomics) against the oracle: histogram and / or q = 0
    curves with 1..3 distinct coverage cutoffs (more fall back to the other kernels), per-item coverage output, item 0,
    short last tiles, <= 3 tail rows, coverage 64 (G = 64: no bit in the one-hot word), and -- with the grid capped to 2
    CTAs -- many tiles per CTA, so that the sixteens words ripple through several shared-memory planes."""
    bits, bitmap, weights = synth.numpy_table(N, G, seed=N * 3 + G)
    if N >= 8:
        bits[1] = 0
        bits[2] = 1
        bits[5] = 1  # full rows: coverage G
        bits[6] = 0
        bits[6, G - 1] = 1
        bitmap = pb.pack_bits(bits)
    if G % 64 == 0:  # a good share of items of the top coverage (64 / 128: no bit in the one-hot words)
        bits[N // 2:N // 2 + N // 7] = 1
        bitmap = pb.pack_bits(bits)
    cases = [[(1, 0.0)], [(3, 0.0)], [(1, 0.0), (2, 0.0)], [(2, 0.0), (1, 0.0), (2, 0.0), (G, 0.0)], [(G + 1, 0.0), (1, 0.0)]]
    all_pairs = sorted({pr for cs in cases for pr in cs})
    exp = oracle_all(bitmap, G, weights, all_pairs)
    dirty = bitmap.copy()
    if G % 64:
        dirty[:, (G - 1) // 64] |= np.uint64(((1 << 64) - 1) ^ ((1 << (G % 64)) - 1))  # garbage above bit G must be ignored
    max_forced = 4 if G <= 64 else 2  # PGX_SCAN_VERT=1 lifts the counter limit for one-word rows only
    with pb.DeviceAbacus(N, G) as a:
        a.upload(dirty, weights)
        for grid in ("2", None):
            if grid:
                monkeypatch.setenv("PGX_SCAN_GRID", grid)
            else:
                monkeypatch.delenv("PGX_SCAN_GRID", raising=False)
            hc, _, ct = a.hist(count=True, weight=False, countable=True)
            assert "k_scan_vert<hist=1,D=0>" in a.last_launch_info(), a.last_launch_info()
            assert np.array_equal(hc, exp["hist"]) and np.array_equal(ct, exp["countable"])
            hc, _, _ = a.hist(count=True, weight=False)
            assert np.array_equal(hc, exp["hist"])
            # the library's own choice: at most two counters (from three on the lane-private kernel is faster);
            # PGX_SCAN_VERT=1 takes the kernel up to its four counters
            for force in (False, True):
                if force:
                    monkeypatch.setenv("PGX_SCAN_VERT", "1")
                else:
                    monkeypatch.delenv("PGX_SCAN_VERT", raising=False)
                for pairs in cases:
                    cov = [c for c, _ in pairs]
                    D = len(set(cov))
                    h2, _, cv = a.hist_ordered_growth(cov, None, weighted=False, hist_count=True, hist_weight=False)
                    assert (f"k_scan_vert<hist=1,D={D}>" in a.last_launch_info()) == (1 + D <= (max_forced if force else 2)), a.last_launch_info()
                    assert np.array_equal(h2, exp["hist"])
                    only = a.ordered_growth(cov, None, weighted=False)
                    assert (f"k_scan_vert<hist=0,D={D}>" in a.last_launch_info()) == (D <= (max_forced if force else 2)), a.last_launch_info()
                    for t, (c, q) in enumerate(pairs):
                        assert np.array_equal(cv[t].astype(np.float64), exp[("node", c, q)]), (grid, pairs, c)
                        assert np.array_equal(only[t].astype(np.float64), exp[("node", c, q)]), (grid, pairs, c)
            monkeypatch.delenv("PGX_SCAN_VERT", raising=False)
            # the weighted modes and more than 3 cutoffs stay on the other kernels
            _, hw, _ = a.hist(count=False, weight=True)
            assert "k_scan_vert" not in a.last_launch_info() and np.array_equal(hw, exp["hist_bp"])


def test_ticket_epilogue_matches(monkeypatch):
    """PGX_SCAN_TICKET=1 keeps the round-1 epilogue (global accumulators, completion ticket, snapshot by the last CTA --
    still what the multi-GPU exchange builds on); the default adds straight into the caller's vector.  Same numbers."""
    monkeypatch.setenv("PGX_SCAN_TICKET", "1")
    for N, G in ((1030, 1024), (4099, 512), (5000, 44)):
        bits, bitmap, weights = synth.numpy_table(N, G, seed=N + 3 * G)
        check_table(bitmap, G, weights)


# ---- node-major -> group-major transpose: both kernels, every tile width ---------------------------------------------

@pytest.mark.parametrize("N,G", [(3000, 20), (40_000, 100), (9000, 200), (33_000, 300), (5000, 700), (20_000, 1024),
                                 (17_000, 1100), (2500, 2100), (70_000, 64)])
def test_transpose_kernels(N, G, monkeypatch):
    """k_transpose_reg (in-register 32x32 transposes, default; tile rows of 2 / 4 / 8 / 16 / 32 32-bit columns, partial last
    column blocks, ragged item counts) and the shuffle-butterfly kernel (PGX_TRANSPOSE=shfl): the group-major copy is
    checked through its consumers -- per-group item counts and the union growth under random group orders, which needs
    every item's bits aligned across all group rows -- against numpy."""
    bits, bitmap, weights = synth.numpy_table(N, G, seed=5 * N + G)
    bits[N // 2] = 1
    bits[N - 1] = 0
    bits[N - 1, G - 1] = 1
    bitmap = pb.pack_bits(bits)
    orders = synth.random_orders(3, G, seed=N + G)
    want_len = bits[1:].sum(axis=0).astype(np.uint64)
    want = [np.maximum.accumulate(bits[1:][:, o], axis=1).sum(axis=0).astype(np.uint64) for o in orders]
    for mode in ("reg", "shfl"):
        if mode == "shfl":
            monkeypatch.setenv("PGX_TRANSPOSE", "shfl")
        else:
            monkeypatch.delenv("PGX_TRANSPOSE", raising=False)
        with pb.DeviceAbacus(N, G) as a:
            a.upload(bitmap, weights)
            _, ln = a.similarity(row_begin=0, row_end=0)
            assert np.array_equal(ln, want_len), mode
            pg = a.permuted_growth(orders, [1], None, weighted=False)
            for k in range(len(orders)):
                assert np.array_equal(pg[k, 0].astype(np.uint64), want[k]), (mode, k)


@pytest.mark.parametrize("N,G", [(30_000, 70), (9000, 300)])
def test_permuted_growth_coverage_sorted_copy(N, G, monkeypatch):
    """k_gm_quorum on the coverage-sorted group-major copy (counting, general thresholds with cutoffs > 1, >= 4 orders):
    warps whose 2048 items all stay below a threshold's coverage cutoff skip its rank logic.  Same curves as on the
    natural copy (PGX_GM_COVSORT=0) and as the oracle, including a cutoff almost no item reaches and one no item does."""
    bits, bitmap, weights = synth.numpy_table(N, G, seed=N + 7 * G)
    orders = synth.random_orders(4, G, seed=5)
    pairs = [(1, 0.0), (2, 0.5), (4, 0.9), (G - 1, 0.3), (G + 1, 0.2), (1, 0.7)]
    cov, thr = cutoffs(G, pairs)
    with pb.DeviceAbacus(N, G) as a:
        a.upload(bitmap, weights)
        monkeypatch.setenv("PGX_GM_COVSORT", "0")
        plain = a.permuted_growth(orders, cov, thr, weighted=False)
        assert "coverage-sorted" not in a.last_launch_info()
        monkeypatch.delenv("PGX_GM_COVSORT")
        got = a.permuted_growth(orders, cov, thr, weighted=False)
        assert "coverage-sorted" in a.last_launch_info(), a.last_launch_info()
        assert np.array_equal(got, plain)
        for p in range(orders.shape[0]):
            exp = oracle_all(pb.pack_bits(bits[:, orders[p]]), G, weights, pairs)
            for t, (c, q) in enumerate(pairs):
                assert np.array_equal(got[p, t].astype(np.float64), exp[("node", c, q)]), (p, c, q)
        # a single order afterwards reuses the copy; the weighted mode keeps its weight-sorted copy
        one = a.permuted_growth(orders[:1], cov, thr, weighted=False)
        assert np.array_equal(one[0], got[0])
        w = a.permuted_growth(orders, cov, thr, weighted=True)
        assert "coverage-sorted" not in a.last_launch_info()
        for t, (c, q) in enumerate(pairs):
            exp = oracle_all(pb.pack_bits(bits[:, orders[1]]), G, weights, pairs)
            assert np.array_equal(w[1, t].astype(np.float64), exp[("bp", c, q)]), (c, q)


def test_arbitrary_cutoff_tables_agree_across_kernels(monkeypatch):
    """The ABI takes a u32 cutoff per position.  The documented tables are ceil((j + 1) q), non-decreasing; k_gm_quorum
    (plane-skipping blocks, cutoffs clamped to j + 2) must also agree with the first-generation k_gm_growth on ANY table
    (cutoffs above j + 1 -- never reachable at position j --, zeros in the middle, huge values), and both with the
    node-major k_scan<quorum>, whose per-word shortcuts assume a non-decreasing table, on non-decreasing ones."""
    N, G = 6000, 300
    bits, bitmap, weights = synth.numpy_table(N, G, seed=77)
    rng = np.random.default_rng(5)
    j = np.arange(G)
    wild = np.stack([rng.integers(0, G + 4, G), j + rng.integers(0, 4, G), np.full(G, 0xFFFFFFFF),
                     np.where(j % 7 == 0, 0, j // 2 + 1)]).astype(np.uint32)
    mono = np.stack([np.cumsum(rng.integers(0, 3, G)), np.ceil((j + 1) * 0.5) + 3, np.full(G, 0xFFFFFFFF),
                     np.minimum(np.cumsum(rng.integers(0, 2, G)) + 1, 40)]).astype(np.uint32)
    wild[0, 0] = 0  # (a table of all zeros would be a q = 0 threshold; none of these is)
    cov = [1, 2, 1, 3]
    orders = synth.random_orders(4, G, seed=8)
    orders[0] = np.arange(G)
    with pb.DeviceAbacus(N, G) as a:
        a.upload(bitmap, weights)
        for thr, with_scan in ((wild, False), (mono, True)):
            for weighted in (False, True):
                monkeypatch.delenv("PGX_GM_QUORUM", raising=False)
                new = a.permuted_growth(orders, cov, thr, weighted=weighted)
                assert "k_gm_quorum" in a.last_launch_info()
                monkeypatch.setenv("PGX_GM_QUORUM", "old")
                old = a.permuted_growth(orders, cov, thr, weighted=weighted)
                monkeypatch.delenv("PGX_GM_QUORUM")
                assert np.array_equal(new, old), (with_scan, weighted)
                assert not new[:, 2].any()  # a cutoff nobody reaches
                if with_scan:
                    monkeypatch.setenv("PGX_QUORUM_PATH", "scan")
                    scan = a.ordered_growth(cov, thr, weighted=weighted)
                    monkeypatch.delenv("PGX_QUORUM_PATH")
                    assert np.array_equal(new[0], scan), weighted


def test_similarity_and_csr_against_independent_witness(monkeypatch):
    """The CUDA path against tests/golden/similarity_csr_witness.json (intersections, group sizes and the CSR {r, c, v}
    counted straight from the GFA's S / P lines, independent of the oracle): CUDA-core and tensor-core similarity
    kernels, bp-weighted similarity, the device CSR."""
    d = json.load(open(os.path.join(GOLDEN, "similarity_csr_witness.json")))
    for case in d["cases"]:
        g, t, op, og, names, bits, weights = fixture_bitmap(case["gfa"], case["count"], groupby_sample=case["grouping"] == "sample",
                                                            groupby_haplotype=case["grouping"] == "haplotype")
        assert list(names) == case["groups"]
        if t.n_items == 0:
            continue
        G = len(names)
        path_group = np.full(len(t.id_prefsum) - 1, -1, dtype=np.int64)
        path_group[op.astype(np.int64)] = og.astype(np.int64)
        with pb.DeviceAbacus(t.n_items, G) as a:
            a.build(t.items, t.id_prefsum, path_group, t.exclude)
            a.upload(None, weights)
            for sim in ("csa", "mma"):
                if sim == "mma" and case["count"] == "bp":
                    continue  # the tensor-core kernel counts; bp sums stay on the CUDA-core kernel
                monkeypatch.setenv("PGX_SIM", sim)
                inter, ln = a.similarity(weighted=(case["count"] == "bp"))
                assert inter.tolist() == case["inter"] and ln.tolist() == case["len"], (case["gfa"], case["grouping"], case["count"], sim)
            monkeypatch.delenv("PGX_SIM")
            if "r" in case:
                r, c, v = a.csr(t.items, t.id_prefsum, path_group, t.exclude)
                assert [int(x) for x in r] == case["r"] and [int(x) for x in c] == case["c"] and [int(x) for x in v] == case["v"], \
                    (case["gfa"], case["grouping"], case["count"])
