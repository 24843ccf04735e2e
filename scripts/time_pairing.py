"""Times tb200_multi_pairing (host-facing) at the sizes of the reference's call sites: t = multi_pairing(comm_list, h_vec)
with 2^10 (C1) and 2^13 (C3) pairs, the sizes of the MIPP rounds in between, and a single pairing (latency floor: one
Miller loop + one final exponentiation). Optional arguments: the Miller team sizes to force (0 = by size, 32, 64, 96)."""
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from oracle import bls12_377 as o  # noqa: E402  (inputs only)
from oracle import bls12_377_g2 as o2  # noqa: E402
from testudo_b200 import _lib, pairing  # noqa: E402

_lib.init()
lib = _lib.engine()
teams = [int(a) for a in sys.argv[1:]] or [0]
base_n = 64
ps, _ = o.rand_points(base_n, 1)
qs, _ = o2.rand_points(base_n, 2)
A = np.array([o.affine_to_words(p) for p in ps], dtype=np.uint64).reshape(-1, 12)
B = np.array([o2.affine_to_words(q) for q in qs], dtype=np.uint64).reshape(-1, 24)
for team in teams:
    lib.tb200_set_pairing_team(team)
    lib.tb200_set_pairing_coop_max(8192)
    for n in (1, 2, 64, 256, 512, 1024, 2048, 4096, 8192):
        reps = (n + base_n - 1) // base_n
        a = np.tile(A, (reps, 1))[:n].copy()
        b = np.tile(B, (reps, 1))[:n].copy()
        pairing.multi_pairing(a, b)
        ts = []
        for _ in range(3):
            t0 = time.perf_counter()
            pairing.multi_pairing(a, b)
            ts.append(time.perf_counter() - t0)
        lib.tb200_set_profiling(1)
        pairing.multi_pairing(a, b)
        st = {k: round(lib.tb200_stage_ms(k.encode()), 3) for k in ("miller", "gt_product", "final_exp")}
        lib.tb200_set_profiling(0)
        print(f"team {team:2d} multi_pairing n={n}: {min(ts) * 1e3:.2f} ms  stages {st}", flush=True)
lib.tb200_set_pairing_team(0)
lib.tb200_set_pairing_coop_max(8192)
