"""Multi-GPU parity tests (need >= 2 CUDA devices; skipped otherwise): the sharded entry points of the C ABI
(pgx_comm_*, pgx_*_sharded, pgx_abacus_broadcast, the fused NVLink exchange) must reproduce the single-GPU results bit
for bit -- one process per GPU under torchrun (tools/check_multigpu.py), one thread per GPU in a single process
(pgx_comm_create_all, the CLI's way), and through `panacus ... --gpus 2`."""
import os
import socket
import subprocess
import sys
import threading

import numpy as np
import pytest

import panacus_b200 as pb
from panacus_b200 import sharding, synth

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
BIN = os.path.join(ROOT, "panacus_b200", "bin", "panacus")


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


@pytest.fixture(autouse=True)
def _need_two_gpus(n_cuda_devices):
    if n_cuda_devices < 2:
        pytest.skip("needs >= 2 CUDA devices")


def test_one_process_per_gpu_under_torchrun(n_cuda_devices):
    n = 2 if n_cuda_devices < 4 else 4
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(ROOT, "tools", "check_multigpu.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert r.returncode == 0 and f"multi-GPU parity ok on {n} GPUs" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]


def test_one_thread_per_gpu_in_one_process(n_cuda_devices):
    world = min(n_cuda_devices, 4)
    N, G = 120_001, 200
    bits, bitmap, weight = synth.numpy_table(N, G, seed=31)
    pairs = [(1, 0.0), (2, 0.5), (3, 0.9)]
    cov = [c for c, _ in pairs]
    thr = np.stack([pb.quorum_thresholds(G, q) for _, q in pairs])
    orders = synth.random_orders(7, G, seed=2)
    with pb.DeviceAbacus(N, G, device=0) as full:
        full.upload(bitmap, weight)
        hc0, hw0, cv0 = full.hist_ordered_growth(cov, thr, weighted=True, hist_count=True, hist_weight=True)
        pg0 = full.permuted_growth(orders, cov, thr)
        in0, ln0 = full.similarity(weighted=True)
        comms = pb.Comm.create_all(list(range(world)))
        reps = [full] + [pb.DeviceAbacus(N, G, device=r) for r in range(1, world)]
        shards = []
        for r in range(world):
            lo, hi = sharding.item_range(N, r, world)
            s = pb.DeviceAbacus(hi - lo, G, device=r)
            s.copy_rows_from(full, lo)
            shards.append(s)
        results, errors = [None] * world, []

        def work(r):
            try:
                reps[r].broadcast(comms[r], root=0, with_weights=True)
                pg = reps[r].permuted_growth_sharded(comms[r], orders, cov, thr)
                inter, ln = reps[r].similarity_sharded(comms[r], weighted=True)
                hc, hw, cv = shards[r].hist_ordered_growth_sharded(comms[r], cov, thr, weighted=True, hist_count=True, hist_weight=True)
                results[r] = (pg, inter, ln, hc, hw, cv)
            except Exception as e:  # noqa: BLE001
                errors.append((r, repr(e)))
        th = [threading.Thread(target=work, args=(r,)) for r in range(world)]
        for t in th:
            t.start()
        for t in th:
            t.join(timeout=600)
        assert not errors, errors
        for r in range(world):
            pg, inter, ln, hc, hw, cv = results[r]
            assert np.array_equal(pg, pg0) and np.array_equal(inter, in0) and np.array_equal(ln, ln0), r
            assert np.array_equal(hc, hc0) and np.array_equal(hw, hw0) and np.array_equal(cv, cv0), r
        for x in shards + reps[1:] + comms:
            x.close()


def _body(text):
    """from the table's first row on (comment lines carry argv; NCCL may log its version when NCCL_DEBUG is set)"""
    lines = text.splitlines()
    k = next(i for i, l in enumerate(lines) if l.startswith("panacus\t") or l.startswith("group\t"))
    return "\n".join(lines[k:])


@pytest.mark.parametrize("sub,extra", [("hist", ["-c", "bp"]), ("hist", ["-c", "node"]),
                                       ("ordered-histgrowth", ["-c", "bp", "-l", "1,2", "-q", "0,0.5"]),
                                       ("similarity", ["-c", "node"]), ("similarity", ["-c", "bp"])])
def test_cli_gpus_flag(sub, extra):
    gfa = os.path.join(GOLDEN, "chrM_test.gfa")
    one = subprocess.run([BIN, sub, gfa, "-S"] + extra, capture_output=True, text=True, timeout=300)
    two = subprocess.run([BIN, sub, gfa, "-S", "--gpus", "2"] + extra, capture_output=True, text=True, timeout=300)
    assert one.returncode == 0 and two.returncode == 0, one.stderr + two.stderr
    assert _body(one.stdout) == _body(two.stdout)
