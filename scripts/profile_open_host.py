"""cProfile of one `Polynomial.open` at 2^NV coefficients (default 26): where the HOST time of the mirror goes."""
import cProfile, pstats, ctypes, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from testudo_b200 import _lib, curve, sqrt_pst, poseidon_transcript
nv = int(sys.argv[1]) if len(sys.argv) > 1 else 26
lib = _lib.engine()
R = curve.R_ORDER
rng = np.random.default_rng(7)
m_col = nv // 2; m_row = nv - m_col; odd = nv % 2
def P(a): return a.ctypes.data_as(ctypes.c_void_p)
def rand_sc(k):
    z = rng.integers(0, 1 << 64, size=(k, 4), dtype=np.uint64); z[:, 3] %= np.uint64(R >> 192); return z
gx = [233578398248691099356572568220835526895379068987715365179118596935057653620464273615301663571204657964920925606294,
      140913150380207355837477652521042157274541796891053068589147167627541651775299824604154852141315666357241556069118,
      63160294768292073209381361943935198908131692476676907196754037919244929611450776219210369229519898517858833747423,
      149157405641012693445398062341192467754805999074082136895788947234480009303640899064710353187729182149407503257491]
g2 = np.array(sum([curve.limbs64(c * curve.FQ_R % curve.Q, 6) for c in gx], []), dtype=np.uint64).reshape(1, 24)
n = 1 << m_row
k = rand_sc(n)
pts2 = np.zeros((n, 24), np.uint64); pts1 = np.zeros((n, 12), np.uint64)
_lib.check(lib.tb200_test_g2_mul(P(np.ascontiguousarray(np.tile(g2, (n, 1)))), P(k), n, P(pts2)))
_lib.check(lib.tb200_test_g1_mul(P(np.ascontiguousarray(np.tile(curve.generator_words().reshape(1, 12), (n, 1)))), P(k), n, P(pts1)))
g_levels = [pts1[: n >> i] for i in range(m_row)]
h_levels = [pts2[: n >> i] for i in range(m_row)]
poly = sqrt_pst.Polynomial.from_evaluations(rand_sc(1 << nv))
ck = sqrt_pst.CommitterKey.from_points(g_levels[0]).with_levels(g_levels, h_levels)
comm_list, t = poly.commit(ck)
point = [int.from_bytes(rng.bytes(40), "little") % R for _ in range(nv)]
def run():
    poly.q = None
    poly.get_q(point)
    tr = poseidon_transcript.PoseidonTranscript("fq")
    return poly.open(tr.as_challenge(), comm_list, ck, point, t)
for _ in range(2): run()
for _ in range(3):
    t0 = time.perf_counter(); run(); print("open incl. get_q: %.1f ms" % ((time.perf_counter() - t0) * 1e3))
pr = cProfile.Profile(); pr.enable(); run(); pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(18)
