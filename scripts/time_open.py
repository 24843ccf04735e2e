"""Where Polynomial::open's wall-clock goes (mirror path), at num_vars = NV (default 26)."""
import ctypes, hashlib, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from testudo_b200 import _lib, curve, fr, mipp, msm, msm_g2, multilinear_pc, sqrt_pst

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
comm_list, _ = poly.commit(ck)
point = [int.from_bytes(rng.bytes(40), "little") % R for _ in range(nv)]
state = hashlib.sha256(b"x")
def challenge(label, points):
    state.update(label)
    for p in points: state.update(np.asarray(p, dtype=np.uint64).tobytes())
    return int.from_bytes(state.digest(), "little") % R or 1
for rep in range(2):
    print("pass", rep, "(the first pass pays one-time allocations / module loads)")
    T = [time.perf_counter()]
    def lap(name):
        T.append(time.perf_counter()); print(f"{name:34s} {(T[-1] - T[-2]) * 1e3:9.2f} ms", flush=True)
    poly.get_q(point); lap("get_q")
    c_u = msm.msm_unchecked(comm_list, poly.chis_b); lap("M2 msm_unchecked(comms, chis)")
    comm_q = sqrt_pst.pc_commit(ck, poly.q); lap("M3 commit(q)")
    pr = mipp.MippProofG1.prove(challenge, comm_list, poly.chis_b, c_u); lap("MIPP G1 only")
    pr = mipp.MippProofG1.prove(challenge, comm_list, poly.chis_b, c_u, h_levels[odd], g_levels[odd:]); lap("MIPP with G2 key + open_g1")
    a_rev = list(point[: m_row])[::-1]
    pf = multilinear_pc.open(h_levels, poly.q, curve.scalars_to_words(a_rev, mont=True)); lap("PST open (G2)")
