import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from testudo_b200 import _lib, curve, sqrt_pst, msm, mipp
from testudo_b200.synthetic import make_bases_dev, make_scalars_dev
_lib.init()
s = int(sys.argv[1]) if len(sys.argv) > 1 else 26
m_row = s - s // 2
z = make_scalars_dev(1 << s, seed=s).cpu().numpy().view(np.uint64)
srs = make_bases_dev(1 << m_row, seed=100 + s).cpu().numpy().view(np.uint64)
ck = sqrt_pst.CommitterKey.from_points(srs)
t0 = time.perf_counter(); pl = sqrt_pst.Polynomial.from_evaluations(z); print("from_evaluations (upload)", time.perf_counter() - t0)
r = [int.from_bytes(np.random.default_rng(s + i).bytes(31), "little") % curve.R_ORDER for i in range(s)]
comm_list, _ = pl.commit(ck)
for rep in range(2):
    pl.q = None
    t0 = time.perf_counter(); pl.get_q(r); t1 = time.perf_counter()
    c_u = msm.msm_unchecked(comm_list, pl.chis_b); t2 = time.perf_counter()
    cq = sqrt_pst.pc_commit(ck, pl.q); t3 = time.perf_counter()
    st = {"k": 5}
    def ch(l, p):
        st["k"] = (st["k"] * 6364136223846793005 + 1442695040888963407) % curve.R_ORDER
        return st["k"] | 1
    pr = mipp.MippProofG1.prove(ch, comm_list, pl.chis_b, c_u); t4 = time.perf_counter()
    print({"get_q": round((t1 - t0) * 1e3, 2), "msm_M2": round((t2 - t1) * 1e3, 2), "commit_q_M3": round((t3 - t2) * 1e3, 2),
           "mipp": round((t4 - t3) * 1e3, 2), "equal": bool(np.array_equal(c_u, cq))})
