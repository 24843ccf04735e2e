"""Wall-clock of every library call inside one `Polynomial.verify` at 2^NV coefficients (default 26), in call order, plus
the host time between calls (transcript replay, scalar arithmetic). The CRS is a real one (trapdoor known to the script),
so the proof verifies."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "benches"))
import pst as bench_pst  # noqa: E402
from testudo_b200 import _lib, curve, sqrt_pst  # noqa: E402
from testudo_b200.poseidon_transcript import PoseidonTranscript  # noqa: E402
from testudo_b200.synthetic import make_scalars_dev  # noqa: E402

nv = int(sys.argv[1]) if len(sys.argv) > 1 else 26
real = _lib.init()
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


m_row = nv - nv // 2
rng = np.random.default_rng(1000 + nv)
z = make_scalars_dev(1 << nv, seed=nv).cpu().numpy().view(np.uint64)
t = [int.from_bytes(rng.bytes(40), "little") % curve.R_ORDER for _ in range(m_row)]
g_levels, h_levels = bench_pst.crs_levels(real, t, False), bench_pst.crs_levels(real, t, True)
ck = sqrt_pst.CommitterKey.from_points(g_levels[0]).with_levels(g_levels, h_levels)
vk = bench_pst.verifier_key(real, t)
pl = sqrt_pst.Polynomial.from_evaluations(z)
r = [int.from_bytes(rng.bytes(40), "little") % curve.R_ORDER for _ in range(nv)]
v = pl.eval(r)
comm_list, t_gt = pl.commit(ck)
opened = pl.open(PoseidonTranscript("fq").as_challenge(), comm_list, ck, r, t_gt)


def run():
    return sqrt_pst.Polynomial.verify(PoseidonTranscript("fq").as_challenge(), vk, opened.u, r, v, opened.pst_proof,
                                      opened.mipp, t_gt)


for _ in range(2):
    assert run() is True
proxy = Timed()
_lib.engine = lambda: proxy
_lib.load = lambda: proxy
LOG.clear()
t_begin = time.perf_counter(); ok = run(); t_end = time.perf_counter()
print(f"verify at 2^{nv}: {(t_end - t_begin) * 1e3:.1f} ms, verdict {ok}, {len(LOG)} library calls")
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
