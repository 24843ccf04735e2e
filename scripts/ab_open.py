"""A/B inside one process: `Polynomial.open` at 2^NV (default 26) with the G2 opening of q started at different points of
the MIPP loop (sqrt_pst.PST_START_LEN: the folded length at which it starts; 2^30 = before the first round)."""
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
lib = _lib.init()
m_row = nv - nv // 2
rng = np.random.default_rng(1000 + nv)
z = make_scalars_dev(1 << nv, seed=nv).cpu().numpy().view(np.uint64)
t = [int.from_bytes(rng.bytes(40), "little") % curve.R_ORDER for _ in range(m_row)]
g_levels, h_levels = bench_pst.crs_levels(lib, t, False), bench_pst.crs_levels(lib, t, True)
ck = sqrt_pst.CommitterKey.from_points(g_levels[0]).with_levels(g_levels, h_levels)
pl = sqrt_pst.Polynomial.from_evaluations(z)
r = [int.from_bytes(rng.bytes(40), "little") % curve.R_ORDER for _ in range(nv)]
comm_list, t_gt = pl.commit(ck)
pl.get_q(r)


def run():
    t0 = time.perf_counter()
    pl.open(PoseidonTranscript("fq").as_challenge(), comm_list, ck, r, t_gt)
    return (time.perf_counter() - t0) * 1e3


run(); run()
for rep in range(2):
    for start in (1 << 30, 4096, 2048, 1024, 512, 256):
        sqrt_pst.PST_START_LEN = start
        ts = [run() for _ in range(4)]
        print(f"PST_START_LEN {start:>10}: best {min(ts):6.2f} ms   all {[round(x, 1) for x in ts]}", flush=True)
