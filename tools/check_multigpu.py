#!/usr/bin/env python3
"""Multi-GPU parity check (run under torchrun on a box with >= 2 GPUs):
  torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29533 tools/check_multigpu.py
Each rank compares the sharded result (item ranges + fused in-kernel NVLink exchange; NCCL variant;
permutations round-robin; similarity row blocks) with the single-GPU result on the whole table."""
import os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import panacus_b200 as pb
from panacus_b200 import sharding, synth

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
N, G = 300_001, 200
bits, bitmap, weight = synth.numpy_table(N, G, seed=77)
pairs = [(1, 0.0), (2, 0.5), (3, 0.9)]
cov = [c for c, _ in pairs]
thr = np.stack([pb.quorum_thresholds(G, q) for _, q in pairs])
with pb.DeviceAbacus(N, G, device=local) as full:
    full.upload(bitmap, weight)
    hc0, hw0, cv0 = full.hist_ordered_growth(cov, thr, weighted=True, hist_count=True, hist_weight=True)
    orders = synth.random_orders(11, G, seed=5)
    pg0 = full.permuted_growth(orders, cov, thr, weighted=False)
    in0, ln0 = full.similarity(weighted=True)
    in1, ln1 = full.similarity(weighted=False)
    pgw0 = full.permuted_growth(orders, cov, thr, weighted=True)
    # work-item sharding behind the C ABI (pgx_comm: NCCL all-gather of device-resident results; bitmap replicated)
    comm = pb.Comm.from_torch_distributed(local)
    pg = full.permuted_growth_sharded(comm, orders, cov, thr, weighted=False)
    assert np.array_equal(pg, pg0), "pgx_permuted_growth_sharded"
    assert np.array_equal(full.permuted_growth_sharded(comm, orders, cov, thr, weighted=True), pgw0), "sharded permuted growth, bp"
    assert np.array_equal(full.permuted_growth_sharded(comm, orders[:1], cov, thr), pg0[:1]), "fewer orders than ranks"
    inter, ln = full.similarity_sharded(comm, weighted=True)
    assert np.array_equal(inter, in0) and np.array_equal(ln, ln0), "pgx_similarity_sharded, bp"
    inter, ln = full.similarity_sharded(comm, weighted=False)
    assert np.array_equal(inter, in1) and np.array_equal(ln, ln1), "pgx_similarity_sharded"
    # replicate rank 0's table over NVLink instead of uploading it on every rank
    with pb.DeviceAbacus(N, G, device=local) as rep:
        if rank == 0:
            rep.upload(bitmap, weight)
        rep.broadcast(comm, root=0, with_weights=True)
        i2, l2 = rep.similarity_sharded(comm, weighted=True)
        assert np.array_equal(i2, in0) and np.array_equal(l2, ln0), "pgx_abacus_broadcast"
    # the torch.distributed plumbing of sharding.py (same partition, host-side gather)
    pg = sharding.sharded_permuted_growth(full, orders, cov, thr, weighted=False, device=dev)
    assert np.array_equal(pg, pg0), "sharded permuted growth"
    inter, ln = sharding.sharded_similarity(full, weighted=True, device=dev)
    assert np.array_equal(inter, in0) and np.array_equal(ln, ln0), "sharded similarity"
bm_r, w_r, n_r = sharding.shard_rows(bitmap, weight, rank, world)
with pb.DeviceAbacus(n_r, G, device=local) as a:
    a.upload(bm_r, w_r)
    # NCCL exchange: behind the C ABI, and the torch.distributed twin
    hc, hw, cv = a.hist_ordered_growth_sharded(comm, cov, thr, weighted=True, hist_count=True, hist_weight=True)
    assert np.array_equal(hc, hc0) and np.array_equal(hw, hw0) and np.array_equal(cv, cv0), "pgx_hist_ordered_growth_sharded"
    hc, hw, cv = sharding.sharded_hist_ordered_growth(a, cov, thr, weighted=True, hist_weight=True, device=dev)
    assert np.array_equal(hc, hc0) and np.array_equal(hw, hw0) and np.array_equal(cv, cv0), "nccl item-range"
    # fused in-kernel exchange, several collective passes in a row (epoch / parity handling); handles via the communicator
    a.exchange_connect_comm(comm)
    for it in range(5):
        hc, hw, cv = a.hist_ordered_growth(cov, thr, weighted=True, hist_count=True, hist_weight=True)
        assert np.array_equal(hc, hc0) and np.array_equal(hw, hw0) and np.array_equal(cv, cv0), ("fused", it)
    # more thresholds than one launch takes (kMaxThresholds = 8): the exchange slots are indexed per launch
    covs = [1 + (t % 3) for t in range(11)]
    many0 = None
    with pb.DeviceAbacus(N, G, device=local) as full2:
        full2.upload(bitmap, weight)
        many0 = full2.ordered_growth(covs, None, weighted=True)
    many = a.ordered_growth(covs, None, weighted=True)
    assert np.array_equal(many, many0), "fused exchange with 11 thresholds"
    h2, _, _ = a.hist()
    assert np.array_equal(h2, hc0)
    a.exchange_disconnect()
    hl, _, _ = a.hist()
    assert int(hl.sum()) == n_r
comm.close()
dist.barrier()
if rank == 0:
    print(f"multi-GPU parity ok on {world} GPUs")
dist.destroy_process_group()
