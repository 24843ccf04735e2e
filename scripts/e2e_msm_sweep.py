"""End-to-end single MSM (N x 2^24 points from ONE host buffer, ONE process, tb200_init_devices) under different upload
schedules: paced vs as-early-as-possible, chunk shapes. Usage: python scripts/e2e_msm_sweep.py [ngpu] [logn]"""
import ctypes, os, sys, time
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from testudo_b200 import _lib, synthetic
ng = int(sys.argv[1]) if len(sys.argv) > 1 else torch.cuda.device_count()
logn = int(sys.argv[2]) if len(sys.argv) > 2 else 24
n = 1 << logn
lib = _lib.init_devices(list(range(ng)))
tot = ng * n
def hostbuf(units, words):
    p = ctypes.c_void_p()
    _lib.check(lib.tb200_host_alloc_sharded(units, words * 8, ctypes.byref(p)))
    return np.ctypeslib.as_array(ctypes.cast(p, ctypes.POINTER(ctypes.c_uint64)), shape=(units, words)), p
B, bp = hostbuf(tot, 12)
S, sp = hostbuf(tot, 4)
b = synthetic.make_bases_dev(n, seed=1); sc = synthetic.make_scalars_dev(n, seed=2)
for r in range(ng):
    torch.from_numpy(B[r * n:(r + 1) * n].view(np.int64)).copy_(b)
    torch.from_numpy(S[r * n:(r + 1) * n].view(np.int64)).copy_(sc)
del b, sc
torch.cuda.empty_cache()
out = np.zeros(12, dtype=np.uint64)
ref = None
def run(name, pace, fr):
    global ref
    arr = (ctypes.c_int * len(fr))(*fr) if fr else None
    _lib.check(lib.tb200_set_host_upload(pace, arr, len(fr)))
    for _ in range(2):
        _lib.check(lib.tb200_msm_g1(bp, sp, tot, 0, out.ctypes.data_as(ctypes.c_void_p)))
    ts = []
    for _ in range(4):
        t0 = time.perf_counter()
        _lib.check(lib.tb200_msm_g1(bp, sp, tot, 0, out.ctypes.data_as(ctypes.c_void_p)))
        ts.append((time.perf_counter() - t0) * 1e3)
    if ref is None: ref = out.copy()
    print(f"{name:52s} min {min(ts):7.2f} ms  all {[round(t, 1) for t in ts]}  same={bool(np.array_equal(out, ref))}", flush=True)
run("built-in 1,2,4,4,3,1,1 as early as possible", 0, [])
run("built-in 1,2,4,4,3,1,1 paced", 1, [])
run("1,1,2,2,2,2,2,2,1,1 paced", 1, [1, 1, 2, 2, 2, 2, 2, 2, 1, 1])
run("16 x 1 paced", 1, [1] * 16)
run("1,1,1,2,2,2,2,2,1,1,1 paced", 1, [1, 1, 1, 2, 2, 2, 2, 2, 1, 1, 1])
run("2,2,2,2,2,2,2,1,1 paced", 1, [2, 2, 2, 2, 2, 2, 2, 1, 1])
run("1,1,2,2,2,2,2,2,1,1 as early as possible", 0, [1, 1, 2, 2, 2, 2, 2, 2, 1, 1])
