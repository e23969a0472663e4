"""Where does a cold process spend its time before the first kernel?  (c5: the CLI's floor.)  Fresh processes only."""
import os, subprocess, sys, time, json

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import ctypes, time, os, sys
t0 = time.perf_counter()
lib = ctypes.CDLL(os.path.join(%r, "panacus_b200", "libpanacus_b200.so"))
t1 = time.perf_counter()
n = ctypes.c_int(0); lib.pgx_device_count(ctypes.byref(n))
t2 = time.perf_counter()
rc = lib.pgx_device_warmup(0)
t3 = time.perf_counter()
h = ctypes.c_void_p()
lib.pgx_abacus_create.argtypes = [ctypes.POINTER(ctypes.c_void_p), ctypes.c_int, ctypes.c_uint64, ctypes.c_uint32]
rc2 = lib.pgx_abacus_create(ctypes.byref(h), 0, 3759736, 44)
t4 = time.perf_counter()
print({"dlopen_ms": (t1-t0)*1e3, "device_count_ms": (t2-t1)*1e3, "warmup_ms": (t3-t2)*1e3, "abacus_create_ms": (t4-t3)*1e3, "ndev": n.value, "rc": [rc, rc2]})
''' % ROOT

def run(env_extra, label):
    env = dict(os.environ); env.update(env_extra)
    for i in range(2):
        t0 = time.perf_counter()
        out = subprocess.run([sys.executable, "-c", CHILD], env=env, capture_output=True, text=True)
        print(label, i, out.stdout.strip(), out.stderr.strip()[-200:], "wall_ms", round((time.perf_counter()-t0)*1e3), flush=True)

print(subprocess.run("nvidia-smi -L; nvidia-smi -q | grep -i -m3 persistence; nproc; free -g | head -2", shell=True, capture_output=True, text=True).stdout)
run({}, "default")
run({"CUDA_VISIBLE_DEVICES": "0"}, "visible0")
run({"CUDA_MODULE_LOADING": "EAGER"}, "eager")
cli = os.path.join(ROOT, "panacus_b200", "bin", "panacus")
gfa = os.path.join(ROOT, "tests", "golden", "chrM_test.gfa")
if os.path.exists(cli) and os.path.exists(gfa):
    for i in range(3):
        t0 = time.perf_counter()
        out = subprocess.run([cli, "histgrowth", gfa, "-c", "bp", "--timing"], capture_output=True, text=True)
        print("cli chrM", i, "wall_ms", round((time.perf_counter()-t0)*1e3), out.stderr.strip()[-400:], flush=True)
