"""Miller-stage time of tb200_multi_pairing in both launch modes (thread-per-pair / CTA-per-pair) over n."""
import sys

import numpy as np

sys.path.insert(0, ".")
from oracle import bls12_377 as o  # noqa: E402  (inputs only)
from oracle import bls12_377_g2 as o2  # noqa: E402
from testudo_b200 import _lib, pairing  # noqa: E402

lib = _lib.init()
ps, _ = o.rand_points(64, 1)
qs, _ = o2.rand_points(64, 2)
A = np.array([o.affine_to_words(p) for p in ps], dtype=np.uint64).reshape(-1, 12)
B = np.array([o2.affine_to_words(q) for q in qs], dtype=np.uint64).reshape(-1, 24)
for n in (256, 512, 1024, 2048, 4096, 8192, 16384):
    a = np.tile(A, (n // 64, 1)).copy()
    b = np.tile(B, (n // 64, 1)).copy()
    row = []
    ref = None
    for mode in (0, 1 << 20):
        lib.tb200_set_pairing_coop_max(mode)
        pairing.multi_pairing(a, b)
        lib.tb200_set_profiling(1)
        out = pairing.multi_pairing(a, b)
        row.append(round(lib.tb200_stage_ms(b"miller"), 3))
        lib.tb200_set_profiling(0)
        assert ref is None or np.array_equal(ref, out)
        ref = out
    print(f"n={n}: miller thread-per-pair {row[0]} ms, CTA-per-pair {row[1]} ms", flush=True)
lib.tb200_set_pairing_coop_max(8192)
