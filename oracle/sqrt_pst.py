"""TEST INFRASTRUCTURE ONLY (never imported by testudo_b200/): the PROVER side of the reference's sqrt-PST commitment
restated with Python integers and the naive group law, for sizes the oracle finishes in seconds (num_vars <= 6).

Follows, line by line:
  * `Polynomial::from_evaluations / get_q / eval / commit / open`   src/sqrt_pst.rs:32-230
  * `MippProof::prove`, `compress`, `compress_field`, `polynomial_evaluations_from_transcript`   src/mipp.rs:31-180,354-383
  * `MultilinearPC::{commit, commit_g2, open, open_g1}`   ark-poly-commit fork (oracle/pst.py)
The transcript is a callback `challenge(label, [(kind, value), ...]) -> int`, kind in {"g1", "g2", "gt"} -- the values
the reference appends at that point (src/mipp.rs:56,97-101,138-141).

PARITY UNPINNED against the arkworks binary: the reference holds no fixtures for these values and cannot be
built here; prover and verifier restatements are pinned against each other (tests/test_oracle_sqrt_pst.py).

ck = dict(nv, powers_of_g[k], powers_of_h[k]) with level k holding the 2^(nv-k) points eq((t_k..), x) * generator.
"""
from __future__ import annotations

from typing import Callable, Dict, List, Sequence

from . import bls12_377 as g1
from . import bls12_377_g2 as g2
from . import pairing as pr
from . import pst

R = g1.R_ORDER


def setup_ck(t: Sequence[int]) -> Dict:
    nv = len(t)
    pg, ph = [], []
    for k in range(nv):
        e = pst.eq_exponents(t[k:])
        pg.append([g1.mul(x, g1.G) for x in e])
        ph.append([g2.mul(x, g2.G2) for x in e])
    return {"nv": nv, "powers_of_g": pg, "powers_of_h": ph}


def get_chi_i(b: Sequence[int], i: int) -> int:
    m = len(b)
    prod = 1
    for j in range(m):
        prod = prod * (b[j] if (i >> (m - j - 1)) & 1 else (1 - b[j])) % R
    return prod


class Polynomial:
    def __init__(self, Z: Sequence[int]):
        n = len(Z)
        num_vars = n.bit_length() - 1
        assert 1 << num_vars == n
        self.m = num_vars // 2
        self.odd = num_vars % 2
        m_col, m_row = self.m, self.m + self.odd
        # polys[i].Z[j] = Z[(j << m_col) | i]   (src/sqrt_pst.rs:48-62)
        self.polys = [[Z[(j << m_col) | i] % R for j in range(1 << m_row)] for i in range(1 << m_col)]
        self.q = None
        self.chis_b = None

    def get_q(self, point: Sequence[int]) -> None:
        b = point[self.m + self.odd:]
        pow_m = 1 << self.m
        chis = [get_chi_i(b, i) for i in range(pow_m)]
        self.q = [sum(self.polys[i][j] * chis[i] for i in range(pow_m)) % R for j in range(pow_m << self.odd)]
        self.chis_b = chis

    def eval(self, point: Sequence[int]) -> int:
        a = point[: len(point) // 2 + self.odd]
        if self.q is None:
            self.get_q(point)
        return sum(qj * get_chi_i(a, j) for j, qj in enumerate(self.q)) % R

    def commit(self, ck: Dict):
        comm_list = [g1.msm_naive(ck["powers_of_g"][0], p) for p in self.polys]
        h_vec = ck["powers_of_h"][self.odd]
        assert len(comm_list) == len(h_vec)
        return comm_list, pr.multi_pairing(comm_list, h_vec)

    def open(self, challenge: Callable, comm_list, ck: Dict, point: Sequence[int], t=None):
        a = list(point[: self.m + self.odd])
        if self.q is None:
            self.get_q(point)
        c_u = g1.msm_naive(comm_list, self.chis_b)
        assert c_u == g1.msm_naive(ck["powers_of_g"][0], self.q)                       # debug_assert, :206
        h_vec = ck["powers_of_h"][self.odd]
        mipp_proof = mipp_prove(challenge, ck, list(comm_list), list(self.chis_b), list(h_vec), c_u, self.odd)
        pst_proof = pst.open_proofs(self.q, a[::-1], ck["powers_of_h"], g2.msm_naive)   # :218-225
        return c_u, pst_proof, mipp_proof


def mipp_prove(challenge: Callable, ck: Dict, a: List, y: List[int], h: List, U, off: int) -> Dict:
    m_a, m_y, m_h = list(a), [v % R for v in y], list(h)
    comms_t, comms_u, xs, xs_inv = [], [], [], []
    challenge(b"U", [("g1", U)])
    while len(m_a) > 1:
        split = len(m_a) // 2
        a_l, a_r = m_a[:split], m_a[split:]
        y_l, y_r = m_y[:split], m_y[split:]
        h_l, h_r = m_h[:split], m_h[split:]
        comm_u_l = g1.msm_naive(a_l, y_r)
        comm_u_r = g1.msm_naive(a_r, y_l)
        comm_t_l = pr.multi_pairing(a_l, h_r)
        comm_t_r = pr.multi_pairing(a_r, h_l)
        c_inv = challenge(b"challenge_i", [("g1", comm_u_l), ("g1", comm_u_r), ("gt", comm_t_l), ("gt", comm_t_r)]) % R
        c = pow(c_inv, -1, R)
        m_a = [g1.add(l, g1.mul(c, r)) for l, r in zip(a_l, a_r)]                      # compress(&mut m_a, split, &c)
        m_y = [(l + r * c_inv) % R for l, r in zip(y_l, y_r)]                          # compress_field
        m_h = [g2.add(l, g2.mul(c_inv, r)) for l, r in zip(h_l, h_r)]                  # compress(&mut m_h, split, &c_inv)
        comms_t.append((comm_t_l, comm_t_r))
        comms_u.append((comm_u_l, comm_u_r))
        xs.append(c)
        xs_inv.append(c_inv)
    final_a, final_h = m_a[0], m_h[0]
    m = len(xs_inv)
    evals = [1]
    for j in range(m):                                                                  # :159-180
        f = xs_inv[m - j - 1]
        evals = evals + [e * f % R for e in evals]
    assert g2.msm_naive(ck["powers_of_h"][off], evals) == final_h                       # debug_assert, :133-134
    rs = [challenge(b"random_point", []) % R for _ in range(m)]
    pst_proof_h = pst.open_proofs(evals, rs, ck["powers_of_g"][off:], g1.msm_naive)      # open_g1, :144
    return {"comms_t": comms_t, "comms_u": comms_u, "final_a": final_a, "final_h": final_h,
            "pst_proof_h": pst_proof_h, "xs_inv": xs_inv, "rs": rs}
