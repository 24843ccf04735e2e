"""TEST INFRASTRUCTURE ONLY (never imported by testudo_b200/): the VERIFIER side of the reference's commitment scheme
with Python integers, so that the reference's own round-trip tests (`check_sqrt_poly_commit`, src/sqrt_pst.rs:297-342:
commit -> open -> verify == true) can be run against the proofs the GPU path produces, with an independent checker.

Restates:
  * `MultilinearPC::check`   (ark-poly-commit 0.4; gadget form src/circuit_verifier.rs:244-312):
        e(C - v g, h) == prod_i e(g_mask[i] - z_i g, proof_i)
  * fork API `check_2`       (gadget form src/circuit_verifier.rs:170-241):
        e(g, C_h - v h) == prod_i e(proof_i, h_mask[off + i] - z_i h),  off = vk.nv - len(z)
  * `MippProof::verify`      src/mipp.rs:182-333
  * `Polynomial::verify`     src/sqrt_pst.rs:232-267
The transcript is the same callback the prover mirror takes: `challenge(label, appended_values) -> int`.

PARITY UNPINNED against the arkworks binary: the reference holds no fixtures for these values and cannot be
built here; prover and verifier restatements are pinned against each other (tests/test_oracle_sqrt_pst.py).

vk = dict(nv, g, h, g_mask[nv], h_mask[nv]) over the oracle's point types (oracle/bls12_377.py, bls12_377_g2.py);
GT values are flat Fq12 tuples (oracle/pairing.py).
"""
from __future__ import annotations

from typing import Callable, Dict, List, Sequence

from . import bls12_377 as g1
from . import bls12_377_g2 as g2
from . import pairing as pr

R = g1.R_ORDER


def setup_vk(t: Sequence[int]) -> Dict:
    """The verifier key of a CRS with trapdoor t (little-endian variables): g_mask[i] = t_i g, h_mask[i] = t_i h."""
    return {"nv": len(t), "g": g1.G, "h": g2.G2, "g_mask": [g1.mul(ti, g1.G) for ti in t],
            "h_mask": [g2.mul(ti, g2.G2) for ti in t]}


def check(vk: Dict, comm: g1.Affine, point: Sequence[int], value: int, proofs: Sequence[g2.Affine2]) -> bool:
    left = pr.pairing(g1.add(comm, g1.neg(g1.mul(value % R, vk["g"]))), vk["h"])
    lefts = [g1.add(vk["g_mask"][i], g1.neg(g1.mul(point[i] % R, vk["g"]))) for i in range(vk["nv"])]
    right = pr.multi_pairing(lefts, list(proofs))
    return left == right


def check_2(vk: Dict, comm_h: g2.Affine2, point: Sequence[int], value: int, proofs: Sequence[g1.Affine]) -> bool:
    off = vk["nv"] - len(point)
    left = pr.pairing(vk["g"], g2.add(comm_h, g2.neg(g2.mul(value % R, vk["h"]))))
    rights = [g2.add(vk["h_mask"][off + i], g2.neg(g2.mul(point[i] % R, vk["h"]))) for i in range(len(point))]
    right = pr.multi_pairing(list(proofs), rights)
    return left == right


def mipp_verify(vk: Dict, challenge: Callable, proof: Dict, point: Sequence[int], U: g1.Affine, T) -> bool:
    """src/mipp.rs:182-333. proof = dict(comms_u [(l, r)], comms_t [(l, r)], final_a, final_h, pst_proof_h [G1]);
    `challenge(label, values)` receives the same VALUES the prover appended, in the test's neutral encoding
    (the caller wraps it so both sides hash identical bytes)."""
    xs, xs_inv = [], []
    final_y = 1
    tc, uc = T, U
    challenge(b"U", [("g1", U)])
    for i, ((ul, ur), (tl, tr)) in enumerate(zip(proof["comms_u"], proof["comms_t"])):
        c_inv = challenge(b"challenge_i", [("g1", ul), ("g1", ur), ("gt", tl), ("gt", tr)]) % R
        c = pow(c_inv, -1, R)
        xs.append(c)
        xs_inv.append(c_inv)
        final_y = final_y * (1 + c_inv * point[i] - point[i]) % R                   # :226
    for (ul, ur), (tl, tr), c, c_inv in zip(proof["comms_u"], proof["comms_t"], xs, xs_inv):
        tc = pr.f12_mul(tc, pr.f12_mul(pr.f12_pow(tl, c_inv), pr.f12_pow(tr, c)))   # :246-265
        uc = g1.add(uc, g1.add(g1.mul(c_inv, ul), g1.mul(c, ur)))
    m = len(xs_inv)
    rs = [challenge(b"random_point", []) % R for _ in range(m)]                     # :281-285
    v = 1
    for i in range(m):
        v = v * (1 + rs[i] * xs_inv[m - i - 1] - rs[i]) % R                         # :294-297
    check_h = check_2(vk, proof["final_h"], rs, v, proof["pst_proof_h"])            # :307
    final_u = g1.mul(final_y, proof["final_a"])                                     # :310
    final_t = pr.pairing(proof["final_a"], proof["final_h"])                        # :311
    return check_h and tc == final_t and uc == final_u


def sqrt_pst_verify(vk: Dict, challenge: Callable, U: g1.Affine, point: Sequence[int], v: int,
                    pst_proof: Sequence[g2.Affine2], mipp_proof: Dict, T) -> bool:
    """src/sqrt_pst.rs:232-267."""
    n = len(point)
    odd = n % 2
    a = list(point[: n // 2 + odd])
    b = list(point[n // 2 + odd:])
    if not mipp_verify(vk, challenge, mipp_proof, b, U, T):
        return False
    return check(vk, U, a[::-1], v, pst_proof)
