"""Latency of small single MSMs through the host-facing call and on the device (CUDA events), Straus path vs pipeline."""
import ctypes, os, sys, time
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from testudo_b200 import _lib, curve, synthetic
lib = _lib.init()
def P(t): return ctypes.c_void_p(t.data_ptr())
for n in (1, 2, 8, 32, 64, 128, 256, 512, 1024, 2048):
    bases = synthetic.make_bases_dev(max(n, 4), seed=3)[:n].contiguous()
    sc = synthetic.make_scalars_dev(n, seed=4)
    out = torch.zeros(12, dtype=torch.int64, device="cuda")
    res = []
    for small in (-1, 0):
        lib.tb200_set_small_msm_max(small)
        for _ in range(3):
            _lib.check(lib.tb200_msm_g1_dev(P(bases), P(sc), n, 0, P(out), None))
        _lib.check(lib.tb200_stream_sync())
        t0 = time.perf_counter()
        reps = 20
        for _ in range(reps):
            _lib.check(lib.tb200_msm_g1_dev(P(bases), P(sc), n, 0, P(out), None))
        _lib.check(lib.tb200_stream_sync())
        res.append((time.perf_counter() - t0) / reps * 1e3)
    lib.tb200_set_small_msm_max(-1)
    print(f"n = {n:5d}: Straus path {res[0]:7.3f} ms   pipeline {res[1]:7.3f} ms", flush=True)
