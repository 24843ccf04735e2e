"""Small deterministic workload for ncu: one single MSM (default 2^20 points, c = 16) run twice."""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from testudo_b200 import _lib  # noqa: E402
from testudo_b200.synthetic import make_bases_dev, make_scalars_dev  # noqa: E402

logn = int(sys.argv[1]) if len(sys.argv) > 1 else 20
c = int(sys.argv[2]) if len(sys.argv) > 2 else 16
lib = _lib.init()
n = 1 << logn
bases = make_bases_dev(n, seed=3)
scal = make_scalars_dev(n, seed=4)
out = torch.zeros(12, dtype=torch.int64, device="cuda")
lib.tb200_set_window_bits(c)
lib.tb200_set_accumulate_mode(int(os.environ.get("TB_MODE", "0")))
for _ in range(2):
    _lib.check(lib.tb200_msm_g1_dev(bases.data_ptr(), scal.data_ptr(), n, 0, out.data_ptr(), None))
    torch.cuda.synchronize()
print("done", out.cpu().numpy()[:2])
