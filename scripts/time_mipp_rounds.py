"""Per-call wall-clock of one MIPP prover run over 2^M points (default M = 13): g1_cross, pairing_cross, g1_fold,
g2_fold for every round."""
import ctypes, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from testudo_b200 import _lib, curve

M = int(sys.argv[1]) if len(sys.argv) > 1 else 13
lib = _lib.engine()
R = curve.R_ORDER
rng = np.random.default_rng(7)
def P(a): return a.ctypes.data_as(ctypes.c_void_p)
def rand_sc(k):
    z = rng.integers(0, 1 << 64, size=(k, 4), dtype=np.uint64); z[:, 3] %= np.uint64(R >> 192); return z
gx = [233578398248691099356572568220835526895379068987715365179118596935057653620464273615301663571204657964920925606294,
      140913150380207355837477652521042157274541796891053068589147167627541651775299824604154852141315666357241556069118,
      63160294768292073209381361943935198908131692476676907196754037919244929611450776219210369229519898517858833747423,
      149157405641012693445398062341192467754805999074082136895788947234480009303640899064710353187729182149407503257491]
g2 = np.array(sum([curve.limbs64(c * curve.FQ_R % curve.Q, 6) for c in gx], []), dtype=np.uint64).reshape(1, 24)
n = 1 << M
k = rand_sc(n)
pts2 = np.zeros((n, 24), np.uint64); pts1 = np.zeros((n, 12), np.uint64)
_lib.check(lib.tb200_test_g2_mul(P(np.ascontiguousarray(np.tile(g2, (n, 1)))), P(k), n, P(pts2)))
_lib.check(lib.tb200_test_g1_mul(P(np.ascontiguousarray(np.tile(curve.generator_words().reshape(1, 12), (n, 1)))), P(k), n, P(pts1)))
y = rand_sc(n)
for rep in range(2):
    ha, hh = ctypes.c_void_p(), ctypes.c_void_p()
    _lib.check(lib.tb200_mipp_g1_begin(P(pts1), P(y), n, 1, ctypes.byref(ha)))
    _lib.check(lib.tb200_mipp_g2_begin(P(pts2), n, 1, ctypes.byref(hh)))
    rows = []
    t_all = time.perf_counter()
    while lib.tb200_mipp_g1_len(ha) > 1:
        ul, ur = np.zeros(12, np.uint64), np.zeros(12, np.uint64)
        tl, tr = np.zeros(72, np.uint64), np.zeros(72, np.uint64)
        c = rand_sc(1)[0]; ci = rand_sc(1)[0]
        t0 = time.perf_counter()
        t1 = time.perf_counter(); _lib.check(lib.tb200_mipp_cross_all(ha, hh, P(ul), P(ur), P(tl), P(tr)))
        t2 = time.perf_counter(); _lib.check(lib.tb200_mipp_g1_fold(ha, P(c), P(ci)))
        t3 = time.perf_counter(); _lib.check(lib.tb200_mipp_g2_fold(hh, P(ci)))
        t4 = time.perf_counter()
        rows.append((lib.tb200_mipp_g1_len(ha) * 2, t1 - t0, t2 - t1, t3 - t2, t4 - t3))
    fh = np.zeros((1, 24), np.uint64)
    t0 = time.perf_counter(); _lib.check(lib.tb200_mipp_g2_read(hh, P(fh))); t1 = time.perf_counter()
    total = time.perf_counter() - t_all
    lib.tb200_mipp_g1_end(ha); lib.tb200_mipp_g2_end(hh)
    print(f"pass {rep}: total {total * 1e3:.1f} ms (final g2 read {(t1 - t0) * 1e3:.2f} ms)")
    print("   len      -      cross_all  g1_fold(enq) g2_fold(enq)   [ms]")
    for r in rows:
        print(f"{r[0]:6d} {r[1] * 1e3:9.2f} {r[2] * 1e3:13.2f} {r[3] * 1e3:9.2f} {r[4] * 1e3:9.2f}")
