"""Cooperative pairing kernels with a 64-lane (two warps, __syncthreads) vs a 32-lane (one warp, __syncwarp) team:
stage times of tb200_multi_pairing at 1 .. 2048 pairs (all CTA-per-pair) and the results' equality."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import bls12_377 as o
from oracle import bls12_377_g2 as o2
from testudo_b200 import _lib, pairing
lib = _lib.init()
ps, _ = o.rand_points(64, 1)
qs, _ = o2.rand_points(64, 2)
A = np.array([o.affine_to_words(p) for p in ps], dtype=np.uint64).reshape(-1, 12)
B = np.array([o2.affine_to_words(q) for q in qs], dtype=np.uint64).reshape(-1, 24)
for n in (1, 2, 64, 256, 1024, 2048):
    a = np.tile(A, ((n + 63) // 64, 1))[:n].copy()
    b = np.tile(B, ((n + 63) // 64, 1))[:n].copy()
    res = {}
    for team in (64, 32):
        lib.tb200_set_pairing_team(team)
        pairing.multi_pairing(a, b)
        lib.tb200_set_profiling(1)
        out = pairing.multi_pairing(a, b)
        st = {k: round(lib.tb200_stage_ms(k.encode()), 3) for k in ("miller", "gt_product", "final_exp")}
        lib.tb200_set_profiling(0)
        res[team] = (out, st)
    print(f"n={n:5d}  team 64 {res[64][1]}  team 32 {res[32][1]}  same={bool(np.array_equal(res[64][0], res[32][0]))}", flush=True)
