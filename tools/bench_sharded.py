#!/usr/bin/env python3
"""BASELINE.json configs[2] and configs[3] on N GPUs (torchrun, one rank per GPU):
  c3: growth under 100 random group orders, 5M items x 512 groups, (coverage, quorum) = (1,0) (2,0.5) (4,0.9);
      order p -> rank p % N, bitmap replicated, NCCL all-gather of the curves
  c4: all-pairs similarity, 10M items x 1024 groups, one block of rows per rank, NCCL all-gather
Each phase: max over ranks of the device-synchronised wall time of the sharded call (includes the exchange).
  torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29544 tools/bench_sharded.py [--quick]"""
import json, os, sys, time
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import panacus_b200 as pb
from panacus_b200 import sharding, synth

quick = "--quick" in sys.argv
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
else:
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:29599", rank=0, world_size=1)

def timed(fn, reps=3):
    out = None
    ts = []
    for _ in range(reps):
        torch.cuda.synchronize(); dist.barrier()
        t0 = time.perf_counter(); out = fn(); torch.cuda.synchronize()
        t = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev if world > 1 else None)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ts.append(float(t.item()))
    return float(np.median(ts)) * 1e3, out

d = dev if world > 1 else None
# ---- c3 ----
N, G, P = (1_000_000, 512, 16) if quick else (5_000_000, 512, 100)
bitmap, weight = synth.torch_table(N, G, seed=synth.SEED_BASE + 3)   # same seed on every rank: replicated bitmap
a = pb.DeviceAbacus(N, G, device=local)
a.adopt_device(bitmap.data_ptr(), weight.data_ptr(), keepalive=(bitmap, weight))
pairs = [(1, 0.0), (2, 0.5), (4, 0.9)]
cov = [c for c, _ in pairs]
thr = np.stack([pb.quorum_thresholds(G, q) for _, q in pairs])
orders = synth.random_orders(P, G, seed=synth.SEED_BASE + 3)
a.permuted_growth(orders[:1], cov, thr)  # builds the group-major copy (not timed)
for weighted in (False, True):
    ms, curves = timed(lambda: sharding.sharded_permuted_growth(a, orders, cov, thr, weighted=weighted, device=d))
    ms0, _ = timed(lambda: sharding.sharded_permuted_growth(a, orders, [1], None, weighted=weighted, device=d))
    if rank == 0:
        chk = int(curves[:, 0, -1].astype(np.uint64).sum() % (1 << 61))
        print(json.dumps({"config": "c3 permuted growth", "n_gpus": world, "N": N, "G": G, "orders": P, "weighted": weighted,
                          "ms_3pairs": round(ms, 3), "ms_q0_only": round(ms0, 3),
                          "cells_per_s_3pairs": N * G * P / ms * 1e3, "checksum": chk}), flush=True)
a.close(); del bitmap, weight
# ---- c4 ----
N, G = (1_000_000, 1024) if quick else (10_000_000, 1024)
bitmap, weight = synth.torch_table(N, G, seed=synth.SEED_BASE + 4)
a = pb.DeviceAbacus(N, G, device=local)
a.adopt_device(bitmap.data_ptr(), weight.data_ptr(), keepalive=(bitmap, weight))
a.similarity(weighted=False, row_begin=0, row_end=1)  # transpose (not timed)
ms, (inter, ln) = timed(lambda: sharding.sharded_similarity(a, weighted=False, device=d))
if rank == 0:
    assert np.array_equal(np.diag(inter), ln) and np.array_equal(inter, inter.T)
    words = (N + 64) // 64
    print(json.dumps({"config": "c4 similarity", "n_gpus": world, "N": N, "G": G, "weighted": False, "ms": round(ms, 3),
                      "pair_words_per_s": G * G * words / ms * 1e3, "checksum": int(inter.sum() % (1 << 61))}), flush=True)
a.close()
dist.barrier()
dist.destroy_process_group()
