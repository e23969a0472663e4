#!/bin/bash
# A/B on one box: k_gm_quorum with 4 (shipped) vs 8 rows in flight per thread (counting, T <= 2): the variant is compiled here
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
run_c3() { timeout 300 python bench.py --workload c3 --steps 10 --warmup 3 --no-e2e > gpurun_out/r2l_c3_$1.json 2> gpurun_out/r2l_c3_$1.err; python - <<PY
import json
d=json.loads(open("gpurun_out/r2l_c3_$1.json").read().strip().splitlines()[-1]); print("c3 $1", d["ms_per_step"], d["roofline"].get("kernel_ms_mean"), d.get("checksum"))
PY
}
run_c3 prefetch4
cd panacus_b200/csrc
cp ../libpanacus_b200.so /tmp/base.so
( time nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC --use_fast_math -DPGX_QUORUM_PREFETCH=8 -c -o /tmp/q8.o pgx_quorum.cu ) 2>&1 | tail -3
nvcc -gencode arch=compute_100a,code=sm_100a -shared -cudart static -o ../libpanacus_b200.so pgx_api.o pgx_scan.o pgx_gm.o pgx_csr.o pgx_comm.o pgx_simmma.o /tmp/q8.o -ldl; echo "link rc=$?"
cd ../..
run_c3 prefetch8
cp /tmp/base.so panacus_b200/libpanacus_b200.so
run_c3 prefetch4_again
