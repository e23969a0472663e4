"""Multi-GPU plumbing for the hot path: one process per GPU, torch.distributed (NCCL on GPUs, gloo in
the CPU tests).  Nothing here computes; it only partitions work and exchanges the integer results.

Two sharding schemes, both exact because every partial result is a u64 sum:

  * item ranges  -- hist / ordered growth: rank r scans items [lo_r, hi_r); one all-reduce (sum) of the
                    KB-sized fused result vector is the path's only exchange step.
  * work items   -- permuted growth: order p goes to rank p % world; similarity: two folded blocks of group rows
                    per rank, each computed from its diagonal rightwards (an equal share of the upper triangle;
                    the matrix is symmetric).  The bitmap is replicated; results are all-gathered.

u64 vectors travel as int64 (same bits; two's-complement wrap-around sums are identical).
"""
from __future__ import annotations

from typing import Optional, Sequence, Tuple

import numpy as np


def item_range(n_items: int, rank: int, world: int) -> Tuple[int, int]:
    """Item ids are 1..=n_items; rank r owns ids [lo, hi) with lo >= 1 (balanced to within one item)."""
    base, rem = divmod(n_items, world)
    lo = 1 + rank * base + min(rank, rem)
    hi = lo + base + (1 if rank < rem else 0)
    return lo, hi


def shard_rows(bitmap: np.ndarray, weight: Optional[np.ndarray], rank: int, world: int):
    """Host-side slice of a node-major table for one rank: the dummy row 0 followed by the rank's items."""
    n_items = bitmap.shape[0] - 1
    lo, hi = item_range(n_items, rank, world)
    bm = np.concatenate([np.zeros((1, bitmap.shape[1]), dtype=bitmap.dtype), bitmap[lo:hi]])
    w = None
    if weight is not None:
        w = np.concatenate([np.zeros(1, dtype=weight.dtype), weight[lo:hi]])
    return bm, w, hi - lo


def order_indices(n_orders: int, rank: int, world: int) -> np.ndarray:
    """Round-robin assignment of permutations: order p -> rank p % world."""
    return np.arange(rank, n_orders, world, dtype=np.int64)


def row_block(n_groups: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous block of similarity rows for one rank."""
    base, rem = divmod(n_groups, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def folded_row_blocks(n_groups: int, rank: int, world: int):
    """Upper-triangle sharding of the similarity matrix: the rows are cut into 2 * world tile-aligned blocks and rank r
    takes blocks r and 2 * world - 1 - r.  A block is computed for the columns >= its first row, so block b costs about
    rows x (G - start_b) and every rank's pair of blocks adds up to the same share.  The boundaries are the ones
    pgx_similarity_sharded uses (pgx_similarity_shard_bounds, host-only)."""
    from .abacus import similarity_shard_bounds
    nb = 2 * world
    bounds = [int(b) for b in similarity_shard_bounds(n_groups, world)]
    return [(bounds[b], bounds[b + 1]) for b in (rank, nb - 1 - rank)]


def _dist():
    import torch.distributed as dist
    return dist


def _to_tensor(arr: np.ndarray, device):
    import torch
    t = torch.from_numpy(np.ascontiguousarray(arr).view(np.int64))
    return t.to(device) if device is not None else t


def allreduce_u64(arr: np.ndarray, device=None, group=None) -> np.ndarray:
    """Element-wise wrapping u64 sum over all ranks (the exchange step of item-range sharding)."""
    dist = _dist()
    t = _to_tensor(arr.astype(np.uint64), device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t.cpu().numpy().view(np.uint64).reshape(arr.shape)


def allgather_rows(local: np.ndarray, counts: Sequence[int], device=None, group=None) -> np.ndarray:
    """Concatenate per-rank row blocks (ragged first dimension) on every rank."""
    import torch
    dist = _dist()
    world = dist.get_world_size(group)
    width = int(np.prod(local.shape[1:])) if local.ndim > 1 else 1
    maxc = max(int(c) for c in counts) if counts else 0
    pad = np.zeros((maxc, width), dtype=np.uint64)
    pad[: local.shape[0]] = local.reshape(local.shape[0], width)
    t = _to_tensor(pad, device)
    outs = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(outs, t, group=group)
    parts = [o.cpu().numpy().view(np.uint64)[: int(c)] for o, c in zip(outs, counts)]
    full = np.concatenate(parts, axis=0) if parts else pad[:0]
    return full.reshape((full.shape[0],) + tuple(local.shape[1:]))


def connect_fused_exchange(abacus, group=None):
    """Wire the in-kernel NVLink all-reduce of `abacus` (a DeviceAbacus holding this rank's item range) to
    the other ranks of `group`: exchange the IPC handles with an all_gather, then map the peers."""
    import torch
    dist = _dist()
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    mine = abacus.exchange_export()
    dev = torch.device("cuda", abacus.device) if dist.get_backend(group) == "nccl" else None
    t = torch.frombuffer(bytearray(mine), dtype=torch.uint8)
    t = t.to(dev) if dev is not None else t
    outs = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(outs, t, group=group)
    handles = b"".join(bytes(o.cpu().numpy().tobytes()) for o in outs)
    abacus.exchange_connect(rank, world, handles)
    dist.barrier(group=group)


# ---- sharded queries (abacus = anything with DeviceAbacus' methods) -----------------------------------------

def sharded_hist_ordered_growth(abacus, cov_abs, quorum_thr=None, weighted=False, hist_weight=False, device=None,
                                group=None, comm=None):
    """Each rank's `abacus` holds its item range.  -> (hist_count, hist_weight | None, curves[T, G]) of the whole
    graph on every rank.  The sum is taken on the curves' first differences == on the curves themselves (linear).
    comm (panacus_b200.Comm): the product path -- pgx_hist_ordered_growth_sharded, ncclAllReduce behind the C ABI."""
    if comm is not None:
        return abacus.hist_ordered_growth_sharded(comm, cov_abs, quorum_thr, weighted=weighted, hist_count=True,
                                                  hist_weight=hist_weight)
    hc, hw, cv = abacus.hist_ordered_growth(cov_abs, quorum_thr, weighted=weighted, hist_count=True,
                                            hist_weight=hist_weight)
    G = hc.shape[0] - 1
    T = cv.shape[0]
    packed = np.concatenate([hc, hw if hw is not None else np.zeros(G + 1, dtype=np.uint64), cv.reshape(-1)])
    total = allreduce_u64(packed, device=device, group=group)
    return total[: G + 1], (total[G + 1: 2 * (G + 1)] if hist_weight else None), total[2 * (G + 1):].reshape(T, G)


def sharded_permuted_growth(abacus, orders: np.ndarray, cov_abs, quorum_thr=None, weighted=False, device=None,
                            group=None, comm=None) -> np.ndarray:
    """Every rank holds the whole bitmap; order p is computed by rank p % world; -> curves [P, T, G] on every rank.
    comm (panacus_b200.Comm): the product path -- pgx_permuted_growth_sharded (device-resident curves, ncclAllGather)."""
    if comm is not None:
        return abacus.permuted_growth_sharded(comm, orders, cov_abs, quorum_thr, weighted=weighted)
    dist = _dist()
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    P = orders.shape[0]
    mine = order_indices(P, rank, world)
    T = len(np.atleast_1d(cov_abs))
    G = orders.shape[1]
    local = abacus.permuted_growth(orders[mine], cov_abs, quorum_thr, weighted=weighted) if len(mine) else \
        np.zeros((0, T, G), dtype=np.uint64)
    counts = [len(order_indices(P, r, world)) for r in range(world)]
    gathered = allgather_rows(local, counts, device=device, group=group)
    out = np.zeros((P, T, G), dtype=np.uint64)
    off = 0
    for r in range(world):
        idx = order_indices(P, r, world)
        out[idx] = gathered[off: off + len(idx)]
        off += len(idx)
    return out


def sharded_similarity(abacus, weighted=False, device=None, group=None, triangle=True, comm=None):
    """Every rank holds the whole bitmap; -> (inter [G, G], len [G]) on every rank.
    triangle=True: two folded row blocks per rank, each from its diagonal rightwards (DeviceAbacus.similarity(upper=True)),
    all-gathered and mirrored -- half the pair work of whole rows.  triangle=False: one block of whole rows per rank.
    comm (panacus_b200.Comm): the product path -- pgx_similarity_sharded (ncclAllGather + device-side assembly); the
    torch.distributed path below is the same partition with host-side plumbing, kept for the CPU (gloo) tests."""
    if comm is not None:
        return abacus.similarity_sharded(comm, weighted=weighted)
    dist = _dist()
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    G = abacus.n_groups
    if not triangle:
        lo, hi = row_block(G, rank, world)
        inter, ln = abacus.similarity(weighted=weighted, row_begin=lo, row_end=hi)
        counts = [row_block(G, r, world)[1] - row_block(G, r, world)[0] for r in range(world)]
        return allgather_rows(inter, counts, device=device, group=group), ln
    import torch
    parts, ln = [], None
    for lo, hi in folded_row_blocks(G, rank, world):
        part, ln = abacus.similarity(weighted=weighted, row_begin=lo, row_end=hi, upper=True)
        parts.append(part)
    counts = [sum(hi - lo for lo, hi in folded_row_blocks(G, r, world)) for r in range(world)]
    maxc = max(counts)
    pad = np.zeros((maxc, G), dtype=np.uint64)
    local = np.concatenate(parts, axis=0)
    pad[: local.shape[0]] = local
    t = _to_tensor(pad, device)
    outs = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(outs, t, group=group)
    # assemble and mirror where the gathered blocks live (the GPU under NCCL): the G x G matrix is touched by a few
    # tensor ops instead of numpy passes over 8 MB per step, then copied to the host once
    upper = torch.zeros((G, G), dtype=torch.int64, device=t.device)
    for r in range(world):
        off = 0
        for lo, hi in folded_row_blocks(G, r, world):
            upper[lo:hi] = outs[r][off: off + hi - lo]
            off += hi - lo
    full = torch.triu(upper) + torch.triu(upper, 1).T
    return full.cpu().numpy().view(np.uint64), ln
