"""CPU-only: the integer core of the general-quorum growth kernel (panacus_b200/csrc/pgx_rank.cuh: bit-sliced rank
counters, mask table, verdict update -- the code k_gm_quorum is instantiated from) replayed on the host by
tests/native/rank_sim.cpp, against the oracle's AbacusByGroup::calc_growth (abacus.rs:989-1032)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from oracle import oracle as po

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "native", "rank_sim.cpp")
HDR = os.path.join(ROOT, "panacus_b200", "csrc", "pgx_rank.cuh")


@pytest.fixture(scope="module")
def sim(tmp_path_factory):
    so = tmp_path_factory.mktemp("rank_sim") / "librank_sim.so"
    subprocess.run(["g++", "-O2", "-std=c++17", "-Wall", "-Wextra", "-x", "c++", "-shared", "-fPIC", "-o", str(so), SRC],
                   check=True)
    L = C.CDLL(str(so))
    vp = C.c_void_p
    L.rank_sim.restype = C.c_int
    L.rank_sim.argtypes = [vp, C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint32, vp, vp, C.c_uint32, vp, vp, C.c_int, vp]
    L.rank_planes_needed_c.restype = C.c_int
    L.rank_planes_needed_c.argtypes = [C.c_uint32]
    return L


def quorum_thr(G, q):
    import math
    return np.array([max(0, int(math.ceil((float(g) + 1.0) * q))) for g in range(G)], dtype=np.uint32)


def group_major(bits):
    """bits u8 [N+1, G] -> u64 [G, ceil((N+1)/64)], bit i%64 of word i/64 of row g"""
    n_rows, G = bits.shape
    nw = (n_rows + 63) // 64
    padded = np.zeros((G, nw * 64), dtype=np.uint8)
    padded[:, :n_rows] = bits.T
    return np.ascontiguousarray(np.packbits(padded, axis=1, bitorder="little").view(np.uint64).reshape(G, nw))


def pack_node_major(bits):
    n_rows, G = bits.shape
    W = (G + 63) // 64
    padded = np.zeros((n_rows, W * 64), dtype=np.uint8)
    padded[:, :G] = bits
    return np.packbits(padded, axis=1, bitorder="little").view(np.uint64).reshape(n_rows, W)


def run_sim(L, bits, order, thr, cov, weight=None, planes=-1):
    n_rows, G = bits.shape
    gm = group_major(bits)
    countable = bits.sum(axis=1).astype(np.uint32)
    countable[0] = 0xFFFFFFFF
    delta = np.zeros(G, dtype=np.int64)
    order = np.ascontiguousarray(order, dtype=np.uint32)
    thr = np.ascontiguousarray(thr, dtype=np.uint32)
    w = None if weight is None else np.ascontiguousarray(weight, dtype=np.uint32)
    rc = L.rank_sim(gm.ctypes.data, gm.shape[1], gm.shape[1], n_rows, G, order.ctypes.data, thr.ctypes.data, int(cov),
                    countable.ctypes.data, None if w is None else w.ctypes.data, planes, delta.ctypes.data)
    assert rc > 0, rc
    return np.cumsum(delta), rc


def oracle_curve(bits, order, cov, q, weight=None):
    n_rows, G = bits.shape
    pb = pack_node_major(bits[:, order])
    items, prefsum, op, og = po.bitmap_to_item_table(pb, G)
    r, c, _ = po.csr_build(n_rows - 1, items, prefsum, op, og)
    if weight is None:
        return po.calc_growth(r, c, G, po.absolute(cov), po.relative(q))
    return po.calc_growth(r, c, G, po.absolute(cov), po.relative(q), count_bp=True, node_lens=weight)


@pytest.mark.parametrize("G", [1, 2, 3, 7, 62, 63, 64, 126, 127, 200, 254, 255, 256, 510, 511, 600])
def test_rank_core_matches_oracle(sim, G):
    rng = np.random.default_rng(G)
    N = 300
    p = rng.random(N + 1)[:, None] ** 2
    bits = (rng.random((N + 1, G)) < p).astype(np.uint8)
    bits[0, :] = 0
    bits[1, :] = 1          # an item in every group: its rank reaches G
    bits[2, :] = 0
    weight = rng.integers(0, 2 ** 32, N + 1, dtype=np.uint64).astype(np.uint32)
    weight[0] = 0
    order = rng.permutation(G).astype(np.uint32)
    for cov, q in ((1, 0.0), (1, 0.1), (2, 0.5), (3, 0.9), (1, 1.0), (2, 0.33)):
        if cov > G:
            continue
        thr = quorum_thr(G, q)
        got, P = run_sim(sim, bits, order, thr, cov)
        assert P == sim.rank_planes_needed_c(G)
        assert np.array_equal(got.astype(np.float64), oracle_curve(bits, order, cov, q)), (G, cov, q)
        gotw, _ = run_sim(sim, bits, order, thr, cov, weight)
        assert np.array_equal(gotw.astype(np.float64), oracle_curve(bits, order, cov, q, weight)), (G, cov, q, "bp")


def test_planes_needed_and_wider_instantiations(sim):
    # the clamp value G + 1 must be representable: 2^P - 1 >= G + 1
    for G, want in ((1, 2), (2, 2), (3, 3), (6, 3), (7, 4), (126, 7), (127, 8), (254, 8), (255, 9), (510, 9), (511, 10),
                    (1022, 10), (1023, 11), (1024, 11), (2 ** 20, 21)):
        assert sim.rank_planes_needed_c(G) == want, G
    rng = np.random.default_rng(3)
    G, N = 100, 200
    bits = (rng.random((N + 1, G)) < 0.5).astype(np.uint8)
    bits[0, :] = 0
    order = np.arange(G, dtype=np.uint32)
    want = oracle_curve(bits, order, 2, 0.5)
    for planes in (7, 8, 9, 10, 11, 12, 14, 16, 21):  # every instantiation the kernel is compiled with
        got, P = run_sim(sim, bits, order, quorum_thr(G, 0.5), 2, planes=planes)
        assert P == planes and np.array_equal(got.astype(np.float64), want)
    # cutoffs beyond G + 1 (never reachable) are clamped, not wrapped
    thr = np.full(G, 2 ** 31, dtype=np.uint32)
    got, _ = run_sim(sim, bits, order, thr, 1)
    assert not got.any()
    thr = np.full(G, G + 1, dtype=np.uint32)
    got, _ = run_sim(sim, bits, order, thr, 1)
    assert not got.any()
