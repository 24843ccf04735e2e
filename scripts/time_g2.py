"""Wall-clock of the sqrt(n)-sized G2 / PST-open pieces at the reference's largest shape (2^13)."""
import ctypes, sys, time, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from testudo_b200 import _lib, curve, msm_g2, msm, multilinear_pc

lib = _lib.engine()
R = curve.R_ORDER
rng = np.random.default_rng(1)
n = 1 << 13
def P(a): return a.ctypes.data_as(ctypes.c_void_p)
def rand_sc(k):
    z = rng.integers(0, 1 << 64, size=(k, 4), dtype=np.uint64); z[:, 3] %= np.uint64(R >> 192); return z
# G2 points: k_i * G2 via the test kernel
gx = [233578398248691099356572568220835526895379068987715365179118596935057653620464273615301663571204657964920925606294,
      140913150380207355837477652521042157274541796891053068589147167627541651775299824604154852141315666357241556069118,
      63160294768292073209381361943935198908131692476676907196754037919244929611450776219210369229519898517858833747423,
      149157405641012693445398062341192467754805999074082136895788947234480009303640899064710353187729182149407503257491]
g2 = np.array(sum([curve.limbs64(c * curve.FQ_R % curve.Q, 6) for c in gx], []), dtype=np.uint64).reshape(1, 24)
def timed(name, fn, reps=3):
    fn(); t0 = time.perf_counter()
    for _ in range(reps): fn()
    print(f"{name:40s} {(time.perf_counter() - t0) / reps * 1e3:9.2f} ms", flush=True)
k = rand_sc(n)
pts2 = np.zeros((n, 24), np.uint64)
G = np.ascontiguousarray(np.tile(g2, (n, 1)))
timed("G2 scalar-mul x 8192 (test kernel)", lambda: _lib.check(lib.tb200_test_g2_mul(P(G), P(k), n, P(pts2))), 1)
pts1 = np.zeros((n, 12), np.uint64)
G1 = np.ascontiguousarray(np.tile(curve.generator_words().reshape(1, 12), (n, 1)))
_lib.check(lib.tb200_test_g1_mul(P(G1), P(k), n, P(pts1)))
sc = rand_sc(n)
for m in (1 << 13, 1 << 10, 1 << 6, 2):
    timed(f"msm_g2 n={m}", lambda: msm_g2.msm_bigint(pts2[:m], sc[:m]))
    timed(f"msm_g1 n={m}", lambda: msm.msm_bigint(pts1[:m], sc[:m]))
timed("compress_g2 split=4096", lambda: msm_g2.compress(pts2, 4096, sc[0], mont=False))
lev2 = [pts2[: n >> i] for i in range(13)]
lev1 = [pts1[: n >> i] for i in range(13)]
ev = rand_sc(n); pt = rand_sc(13)
timed("pst_open_g2 nv=13", lambda: multilinear_pc.open(lev2, ev, pt, mont=True))
timed("pst_open_g1 nv=13", lambda: multilinear_pc.open_g1(lev1, ev, pt, mont=True))
