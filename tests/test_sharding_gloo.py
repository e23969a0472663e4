"""world_size-2 (and 3) gloo tests of the multi-GPU plumbing (panacus_b200/sharding.py) on CPU.
The per-rank compute is a numpy stand-in with DeviceAbacus' method signatures (test double only); what is
under test is the partitioning and the exchange: item-range all-reduce, order round-robin + all-gather,
similarity row blocks + all-gather must reproduce the single-process result exactly."""
import os
import socket

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

from panacus_b200 import quorum_thresholds, sharding, synth


class NumpyAbacus:
    """Test double: same interface as panacus_b200.DeviceAbacus, dense numpy arithmetic."""

    def __init__(self, bits, weight):
        self.bits = bits.astype(np.int64)
        self.bits[0] = 0
        self.w = weight.astype(np.int64)
        self.n_items, self.n_groups = bits.shape[0] - 1, bits.shape[1]

    def _curve(self, bits, c, thr, weighted):
        G = self.n_groups
        total = bits.sum(1)
        cnt = np.cumsum(bits, 1)
        col = np.arange(G)[None, :]
        last = np.maximum.accumulate(np.where(bits > 0, col, -1), axis=1)
        ok = (last >= 0) & (cnt >= thr[np.clip(last, 0, G - 1)]) & (total[:, None] >= c)
        w = self.w if weighted else np.ones_like(self.w)
        return (ok * w[:, None]).sum(0).astype(np.uint64)

    def hist_ordered_growth(self, cov_abs, quorum_thr=None, weighted=False, hist_count=True, hist_weight=False):
        G = self.n_groups
        cov = self.bits.sum(1)
        hc = np.bincount(cov[1:], minlength=G + 1).astype(np.uint64)
        hw = np.bincount(cov[1:], weights=self.w[1:], minlength=G + 1).astype(np.uint64) if hist_weight else None
        T = len(cov_abs)
        thr = np.zeros((T, G), dtype=np.int64) if quorum_thr is None else np.asarray(quorum_thr, dtype=np.int64).reshape(T, G)
        cv = np.stack([self._curve(self.bits, cov_abs[t], thr[t], weighted) for t in range(T)])
        return hc, hw, cv

    def permuted_growth(self, orders, cov_abs, quorum_thr=None, weighted=False):
        T, G = len(cov_abs), self.n_groups
        thr = np.zeros((T, G), dtype=np.int64) if quorum_thr is None else np.asarray(quorum_thr, dtype=np.int64).reshape(T, G)
        return np.stack([np.stack([self._curve(self.bits[:, o], cov_abs[t], thr[t], weighted) for t in range(T)])
                         for o in orders]) if len(orders) else np.zeros((0, T, G), dtype=np.uint64)

    def similarity(self, weighted=False, row_begin=0, row_end=None, upper=False):
        w = self.w if weighted else np.ones_like(self.w)
        inter = (self.bits.T @ (self.bits * w[:, None])).astype(np.uint64)
        part = inter[row_begin:row_end].copy()
        if upper:
            part[:, :row_begin] = 0  # pgx_similarity_upper: columns left of the block's first row are not computed
        return part, np.diag(inter).copy()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        N, G = 1501, 70
        bits, bitmap, weight = synth.numpy_table(N, G, seed=123)
        full = NumpyAbacus(bits, weight)
        cov = [1, 2, 3]
        thr = np.stack([quorum_thresholds(G, q) for q in (0.0, 0.5, 0.9)])
        # --- item-range sharding + all-reduce
        lo, hi = sharding.item_range(N, rank, world)
        local_bits = np.concatenate([np.zeros((1, G), dtype=bits.dtype), bits[lo:hi]])
        local_w = np.concatenate([np.zeros(1, dtype=weight.dtype), weight[lo:hi]])
        bm_r, w_r, n_r = sharding.shard_rows(bitmap, weight, rank, world)
        assert n_r == hi - lo and np.array_equal(w_r, local_w) and bm_r.shape[0] == n_r + 1
        local = NumpyAbacus(local_bits, local_w)
        hc, hw, cv = sharding.sharded_hist_ordered_growth(local, cov, thr, weighted=True, hist_weight=True)
        hc0, hw0, cv0 = full.hist_ordered_growth(cov, thr, weighted=True, hist_weight=True)
        assert np.array_equal(hc, hc0) and np.array_equal(hw, hw0) and np.array_equal(cv, cv0)
        # u64 wrap-around survives the int64 transport
        big = np.array([np.uint64(0xFFFFFFFFFFFFFFFF), np.uint64(1) << np.uint64(63)], dtype=np.uint64)
        s = sharding.allreduce_u64(big)
        want = (big.astype(object) * world) % (1 << 64)
        assert [int(x) for x in s] == [int(x) for x in want]
        # --- permutations round-robin + all-gather (5 orders over `world` ranks: ragged)
        orders = synth.random_orders(5, G, seed=9)
        pg = sharding.sharded_permuted_growth(full, orders, cov, thr, weighted=False)
        assert np.array_equal(pg, full.permuted_growth(orders, cov, thr, weighted=False))
        # --- similarity: folded upper-triangle blocks (default) and whole-row blocks, + all-gather
        inter0, ln0 = full.similarity(weighted=True)
        for triangle in (True, False):
            inter, ln = sharding.sharded_similarity(full, weighted=True, triangle=triangle)
            assert np.array_equal(inter, inter0) and np.array_equal(ln, ln0), triangle
        # a wider table: several 64-row tiles per block boundary
        bits2, _, weight2 = synth.numpy_table(301, 333, seed=5)
        wide = NumpyAbacus(bits2, weight2)
        inter1, ln1 = wide.similarity(weighted=False)
        inter, ln = sharding.sharded_similarity(wide, weighted=False, triangle=True)
        assert np.array_equal(inter, inter1) and np.array_equal(ln, ln1)
        ret[rank] = "ok"
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharding_over_gloo(world):
    port = _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    assert dict(ret) == {r: "ok" for r in range(world)}


def test_similarity_shard_bounds():
    """pgx_similarity_shard_bounds (host-only entry point of the C ABI): the 2 * world row blocks of the sharded
    similarity cover the rows once, are 64-row tile aligned, and every rank's folded pair carries the same share of the
    upper triangle's tile work when the tiles divide evenly."""
    from panacus_b200 import similarity_shard_bounds
    for G, world in [(1024, 8), (1024, 2), (1024, 1), (512, 4), (100, 4), (44, 8), (1, 1), (65, 2), (1000, 3), (4096, 8)]:
        b = [int(x) for x in similarity_shard_bounds(G, world)]
        assert len(b) == 2 * world + 1 and b[0] == 0 and b[-1] == G
        assert all(b[k] <= b[k + 1] for k in range(2 * world))
        assert all(x % 64 == 0 for x in b[:-1])
        blocks = [blk for r in range(world) for blk in sharding.folded_row_blocks(G, r, world)]
        rows = sorted(x for lo, hi in blocks for x in range(lo, hi))
        assert rows == list(range(G))
        tiles = (G + 63) // 64
        if tiles % (2 * world) == 0:  # even split: equal tile work per rank (rows x columns right of the block start)
            work = [sum(((hi - lo) // 64) * (tiles - lo // 64) for lo, hi in sharding.folded_row_blocks(G, r, world))
                    for r in range(world)]
            assert len(set(work)) == 1, (G, world, work)


def test_partitions_cover_everything_once():
    for n, w in [(10, 3), (7, 8), (1000, 8), (1, 2)]:
        ids = []
        for r in range(w):
            lo, hi = sharding.item_range(n, r, w)
            ids += list(range(lo, hi))
        assert ids == list(range(1, n + 1))
        rows = []
        for r in range(w):
            lo, hi = sharding.row_block(n, r, w)
            rows += list(range(lo, hi))
        assert rows == list(range(n))
        assert sorted(np.concatenate([sharding.order_indices(n, r, w) for r in range(w)]).tolist()) == list(range(n))
