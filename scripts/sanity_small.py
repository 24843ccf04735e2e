"""Small end-to-end exercise of every kernel family (for compute-sanitizer): single MSM (both XYZZ kernels and the
affine rounds), batch with row sort, MIPP fold, get_q."""
import ctypes, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from testudo_b200 import _lib, msm, sqrt_pst, mipp
from testudo_b200.synthetic import make_bases_dev, make_scalars_dev, expected_msm
lib = _lib.init()
n = 1 << 13
bases = make_bases_dev(n, seed=1); scal = make_scalars_dev(n, seed=2, skew=True)
exp = expected_msm(scal, n, seed=1)
B = bases.cpu().numpy().view(np.uint64); S = scal.cpu().numpy().view(np.uint64)
for mode in (1, 2, 3):
    lib.tb200_set_accumulate_mode(mode)
    assert np.array_equal(msm.msm_bigint(B, S), exp), mode
lib.tb200_set_accumulate_mode(0)
srs = make_bases_dev(64, seed=3).cpu().numpy().view(np.uint64)
ck = sqrt_pst.CommitterKey.from_points(srs)
z = make_scalars_dev(256 * 64, seed=4).cpu().numpy().view(np.uint64)
out = np.zeros((256, 12), dtype=np.uint64)
_lib.check(lib.tb200_msm_g1_batch(ck._h, z.ctypes.data_as(ctypes.c_void_p), 256, 64, 1, 256, 0, out.ctypes.data_as(ctypes.c_void_p)))
v = mipp.compress(B[:64], 32, S[5])
poly = sqrt_pst.Polynomial.from_evaluations(z[:1 << 12])
poly.get_q([3 + i for i in range(12)])
print("sanity ok")
