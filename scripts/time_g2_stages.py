"""Stage split (CUDA events inside the library) of single G2 MSMs of 2^13 .. 2 points."""
import ctypes, sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from testudo_b200 import _lib, curve, msm_g2
lib = _lib.engine()
R = curve.R_ORDER
rng = np.random.default_rng(1)
n = 1 << 13
def P(a): return a.ctypes.data_as(ctypes.c_void_p)
def rand_sc(k):
    z = rng.integers(0, 1 << 64, size=(k, 4), dtype=np.uint64); z[:, 3] %= np.uint64(R >> 192); return z
gx = [233578398248691099356572568220835526895379068987715365179118596935057653620464273615301663571204657964920925606294,
      140913150380207355837477652521042157274541796891053068589147167627541651775299824604154852141315666357241556069118,
      63160294768292073209381361943935198908131692476676907196754037919244929611450776219210369229519898517858833747423,
      149157405641012693445398062341192467754805999074082136895788947234480009303640899064710353187729182149407503257491]
g2 = np.array(sum([curve.limbs64(c * curve.FQ_R % curve.Q, 6) for c in gx], []), dtype=np.uint64).reshape(1, 24)
k = rand_sc(n)
pts2 = np.zeros((n, 24), np.uint64)
_lib.check(lib.tb200_test_g2_mul(P(np.ascontiguousarray(np.tile(g2, (n, 1)))), P(k), n, P(pts2)))
sc = rand_sc(n)
lib.tb200_set_profiling(1)
for m in (1 << 13, 1 << 10, 1 << 6, 2):
    for _ in range(2): msm_g2.msm_bigint(pts2[:m], sc[:m])
    st = {s: round(lib.tb200_stage_ms(s.encode()), 3) for s in ("digits", "scan", "scatter", "accumulate", "fixup", "reduce", "finalize", "total")}
    c = ctypes.c_int(); W = ctypes.c_int(); K = ctypes.c_int(); M = ctypes.c_uint64(); B = ctypes.c_uint64()
    lib.tb200_last_geometry(ctypes.byref(c), ctypes.byref(W), ctypes.byref(M), ctypes.byref(B), ctypes.byref(K))
    print(f"msm_g2 n={m}: c={c.value} W={W.value} K={K.value} {st}", flush=True)
