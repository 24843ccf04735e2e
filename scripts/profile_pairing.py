"""ncu target: one multi_pairing of n pairs (argv[1], default 64) after a warm-up call."""
import sys

import numpy as np

sys.path.insert(0, ".")
from oracle import bls12_377 as o  # noqa: E402  (inputs only)
from oracle import bls12_377_g2 as o2  # noqa: E402
from testudo_b200 import _lib, pairing  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
_lib.init()
ps, _ = o.rand_points(64, 1)
qs, _ = o2.rand_points(64, 2)
A = np.array([o.affine_to_words(p) for p in ps], dtype=np.uint64).reshape(-1, 12)
B = np.array([o2.affine_to_words(q) for q in qs], dtype=np.uint64).reshape(-1, 24)
reps = (n + 63) // 64
a = np.tile(A, (reps, 1))[:n].copy()
b = np.tile(B, (reps, 1))[:n].copy()
pairing.multi_pairing(a, b)
pairing.multi_pairing(a, b)
