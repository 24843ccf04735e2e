"""Wall-clock of every library call inside one `Polynomial.open` at 2^NV coefficients (default 26), in call order, plus the
time spent between calls (host code of the mirror: transcript, conversions)."""
import ctypes, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from testudo_b200 import _lib, curve, sqrt_pst, poseidon_transcript
nv = int(sys.argv[1]) if len(sys.argv) > 1 else 26
real = _lib.engine()
LOG = []


class Timed:
    def __getattr__(self, name):
        fn = getattr(real, name)
        if not name.startswith("tb200_"):
            return fn

        def call(*a):
            t0 = time.perf_counter()
            r = fn(*a)
            LOG.append((name, t0, time.perf_counter()))
            return r
        return call


proxy = Timed()
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
_lib.check(real.tb200_test_g2_mul(P(np.ascontiguousarray(np.tile(g2, (n, 1)))), P(k), n, P(pts2)))
_lib.check(real.tb200_test_g1_mul(P(np.ascontiguousarray(np.tile(curve.generator_words().reshape(1, 12), (n, 1)))), P(k), n, P(pts1)))
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
_lib.engine = lambda: proxy      # every module calls _lib.engine() / _lib.load() per call
_lib.load = lambda: proxy
LOG.clear()
t_begin = time.perf_counter(); run(); t_end = time.perf_counter()
print(f"open incl. get_q: {(t_end - t_begin) * 1e3:.1f} ms, {len(LOG)} library calls")
agg = {}
prev = t_begin
gaps = 0.0
for name, a, b in LOG:
    agg.setdefault(name, [0, 0.0]); agg[name][0] += 1; agg[name][1] += b - a
    gaps += a - prev; prev = b
gaps += t_end - prev
for name, (cnt, tot) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{name:36s} x{cnt:4d} {tot * 1e3:8.2f} ms")
print(f"{'host code between calls':36s}       {gaps * 1e3:8.2f} ms")
print("-- in order (calls longer than 0.3 ms or gaps longer than 0.3 ms) --")
prev = t_begin
for name, a, b in LOG:
    if a - prev > 0.3e-3: print(f"   [host {(a - prev) * 1e3:6.2f} ms]")
    if b - a > 0.3e-3: print(f"{(a - t_begin) * 1e3:8.2f}  {name:34s} {(b - a) * 1e3:7.2f} ms")
    prev = b
