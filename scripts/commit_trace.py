"""tb200_msm_g1_batch_ptrs / tb200_sqrt_pst_commit from host rows on N GPUs with TB200_TRACE=1: where the call's wall time goes."""
import ctypes, os, sys, time
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from testudo_b200 import _lib, synthetic
ng = int(sys.argv[1]) if len(sys.argv) > 1 else torch.cuda.device_count()
lib = _lib.init_devices(list(range(ng)))
rows = cols = 1 << 13
srs = synthetic.make_bases_dev(cols, seed=777).cpu().numpy().view(np.uint64)
h = ctypes.c_void_p(); _lib.check(lib.tb200_srs_load(srs.ctypes.data_as(ctypes.c_void_p), cols, 0, ctypes.byref(h)))
p = ctypes.c_void_p(); _lib.check(lib.tb200_host_alloc_sharded(rows, cols * 32, ctypes.byref(p)))
Z = np.ctypeslib.as_array(ctypes.cast(p, ctypes.POINTER(ctypes.c_uint64)), shape=(rows, cols * 4))
for r0 in range(0, rows, 512):
    zz = synthetic.make_scalars_dev(512 * cols, seed=r0)
    torch.from_numpy(Z[r0:r0 + 512].view(np.int64)).copy_(zz.view(512, cols * 4))
ptrs = (ctypes.c_void_p * rows)(*[p.value + i * cols * 32 for i in range(rows)])
out = np.zeros((rows, 12), dtype=np.uint64)
for mont in (1, 0):
    for rep in range(3):
        t0 = time.perf_counter()
        _lib.check(lib.tb200_msm_g1_batch_ptrs(h, ptrs, rows, cols, mont, out.ctypes.data_as(ctypes.c_void_p)))
        print(f"mont={mont} rep {rep}: {(time.perf_counter() - t0) * 1e3:.2f} ms", file=sys.stderr, flush=True)
