"""ncu target: the verifier's three device steps at the sizes of a 2^26 proof (m = 13 rounds) after a warm-up pass --
tb200_gt_multi_pow (27 powers + product), tb200_msm_g1_each (14 rows of 2 points), tb200_multi_pairing_batch (5 products
of up to 14 pairs)."""
import sys

import numpy as np

sys.path.insert(0, ".")
from oracle import bls12_377 as o  # noqa: E402  (inputs only)
from oracle import bls12_377_g2 as o2  # noqa: E402
from testudo_b200 import _lib, msm, pairing  # noqa: E402

_lib.init()
ps, _ = o.rand_points(28, 1)
qs, _ = o2.rand_points(14, 2)
A = np.array([o.affine_to_words(p) for p in ps], dtype=np.uint64).reshape(-1, 12)
B = np.array([o2.affine_to_words(q) for q in qs], dtype=np.uint64).reshape(-1, 24)
ks = np.array([o.to_limbs64(k, 4) for k in o.rand_scalars(28, 3)], dtype=np.uint64)
gt = np.stack([pairing.pairing(A[i], B[i % 14]) for i in range(3)])
bases = np.tile(gt, (9, 1))
for _ in range(2):
    pairing.gt_multi_pow(bases, ks[:27])
    msm.msm_each(A, ks, 2)
    pairing.multi_pairing_batch([(A[:1], B[:1]), (A[1:2], B[1:2]), (A[:14], B), (A[2:3], B[2:3]), (A[14:27], B[:13])])
