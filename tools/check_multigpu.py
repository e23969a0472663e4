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
    # work-item sharding over NCCL (bitmap replicated)
    pg = sharding.sharded_permuted_growth(full, orders, cov, thr, weighted=False, device=dev)
    assert np.array_equal(pg, pg0), "sharded permuted growth"
    inter, ln = sharding.sharded_similarity(full, weighted=True, device=dev)
    assert np.array_equal(inter, in0) and np.array_equal(ln, ln0), "sharded similarity"
bm_r, w_r, n_r = sharding.shard_rows(bitmap, weight, rank, world)
with pb.DeviceAbacus(n_r, G, device=local) as a:
    a.upload(bm_r, w_r)
    # NCCL exchange
    hc, hw, cv = sharding.sharded_hist_ordered_growth(a, cov, thr, weighted=True, hist_weight=True, device=dev)
    assert np.array_equal(hc, hc0) and np.array_equal(hw, hw0) and np.array_equal(cv, cv0), "nccl item-range"
    # fused in-kernel exchange, several collective passes in a row (epoch / parity handling)
    sharding.connect_fused_exchange(a)
    for it in range(5):
        hc, hw, cv = a.hist_ordered_growth(cov, thr, weighted=True, hist_count=True, hist_weight=True)
        assert np.array_equal(hc, hc0) and np.array_equal(hw, hw0) and np.array_equal(cv, cv0), ("fused", it)
    h2, _, _ = a.hist()
    assert np.array_equal(h2, hc0)
    a.exchange_disconnect()
    hl, _, _ = a.hist()
    assert int(hl.sum()) == n_r
dist.barrier()
if rank == 0:
    print(f"multi-GPU parity ok on {world} GPUs")
dist.destroy_process_group()
