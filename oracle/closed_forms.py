"""TEST INFRASTRUCTURE ONLY (never imported by testudo_b200/): closed forms over KNOWN DISCRETE LOGARITHMS, so that the
prover values can be checked at the reference's full sizes (BASELINE.json configs[1..3]) where the naive big-integer
restatements (oracle/sqrt_pst.py) would take hours.

If a_i = alpha_i G, h_i = eta_i G2 and y_i are scalars, every value `MippProof::prove` emits (src/mipp.rs:58-122) is a
group element whose discrete log is an inner product mod r:
    comm_u_l = MSM(a_l, y_r)              = (sum_i alpha_i y_{s+i}) G                       src/mipp.rs:77-85
    comm_u_r = MSM(a_r, y_l)              = (sum_i alpha_{s+i} y_i) G
    comm_t_l = prod_i e(a_i, h_{s+i})     = e(G, G2)^(sum_i alpha_i eta_{s+i})               src/mipp.rs:87-94
    comm_t_r = prod_i e(a_{s+i}, h_i)     = e(G, G2)^(sum_i alpha_{s+i} eta_i)
    fold:  alpha <- alpha_l + c alpha_r,  y <- y_l + c_inv y_r,  eta <- eta_l + c_inv eta_r   src/mipp.rs:106-114
The group elements themselves come from the oracle's own scalar multiplications (bls12_377.mul, bls12_377_g2.mul) and
GT powers (pairing.f12_pow of the oracle's e(G, G2)): nothing here touches the engine.

PARITY UNPINNED against the arkworks binary like the rest of oracle/ (DESIGN.md 2): these are identities of the
definitions, not reference outputs.
"""
from __future__ import annotations

from typing import Dict, List, Sequence

from . import bls12_377 as g1
from . import bls12_377_g2 as g2
from . import pairing as pr

R = g1.R_ORDER


def mipp_dlogs(alphas: Sequence[int], etas: Sequence[int], ys: Sequence[int], c_invs: Sequence[int]) -> Dict:
    """Discrete logs of every MIPP prover value for the given per-round challenges (c_inv as the transcript returns it,
    src/mipp.rs:97-106). Returns dict(u=[(l, r)], t=[(l, r)], final_a, final_y, final_h)."""
    a, e, y = [v % R for v in alphas], [v % R for v in etas], [v % R for v in ys]
    assert len(a) == len(e) == len(y) and len(a) == 1 << len(c_invs)
    us, ts = [], []
    for c_inv in c_invs:
        s = len(a) // 2
        c = pow(c_inv, -1, R)
        us.append((sum(a[i] * y[s + i] for i in range(s)) % R, sum(a[s + i] * y[i] for i in range(s)) % R))
        ts.append((sum(a[i] * e[s + i] for i in range(s)) % R, sum(a[s + i] * e[i] for i in range(s)) % R))
        a = [(a[i] + c * a[s + i]) % R for i in range(s)]
        y = [(y[i] + c_inv * y[s + i]) % R for i in range(s)]
        e = [(e[i] + c_inv * e[s + i]) % R for i in range(s)]
    return {"u": us, "t": ts, "final_a": a[0], "final_y": y[0], "final_h": e[0]}


_E = None


def gt_generator():
    """e(G, G2) by the oracle's pairing (cached)."""
    global _E
    if _E is None:
        _E = pr.pairing(g1.G, g2.G2)
    return _E


def g1_of(dlog: int):
    return g1.mul(dlog % R, g1.G) if dlog % R else None


def g2_of(dlog: int):
    return g2.mul(dlog % R, g2.G2) if dlog % R else None


def gt_of(dlog: int):
    return pr.f12_pow(gt_generator(), dlog % R)


def row_dlogs(z_rows: Sequence[Sequence[int]], srs_dlogs: Sequence[int]) -> List[int]:
    """dlog of MSM(srs, row) for each row of canonical scalars: the row commitments of `Polynomial::commit`
    (src/sqrt_pst.rs:121-125) over an SRS with known discrete logs."""
    return [sum(int(z) * d for z, d in zip(row, srs_dlogs)) % R for row in z_rows]
