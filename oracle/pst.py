"""TEST INFRASTRUCTURE ONLY: restatement of ark-poly-commit 0.4 `multilinear_pc::MultilinearPC::{open, open_g1}`
(un-vendored git dependency of the reference, Cargo.toml:34,73; behaviour restated in SURVEY.md App. A.2/A.3; call
sites src/sqrt_pst.rs:225 and src/mipp.rs:144) with Python integers, generic over the group.

    for i in 0..nv, k = nv - i:
        q_k[b]     = r_k[2b+1] - r_k[2b]
        r_{k-1}[b] = r_k[2b] * (1 - point[i]) + r_k[2b+1] * point[i]
        proof_i    = MSM(level_bases[i], [q_k[x >> 1] for x in 0..2^k])

Also the synthetic CRS `powers[k][x] = eq((t_k..t_{nv-1}), x) * G` (little-endian variables) as exponents, which
gives the closed form proof_i = q_k(t_{i+1}, .., t_{nv-1}) * G used at sizes the naive MSM cannot reach.
"""
from __future__ import annotations

from typing import Callable, List, Sequence

from . import bls12_377 as g1

R = g1.R_ORDER


def quotients(evals: Sequence[int], point: Sequence[int]) -> List[List[int]]:
    """[q_nv, q_{nv-1}, .., q_1] in the order the proofs are produced (i = 0..nv-1)."""
    nv = len(point)
    assert len(evals) == 1 << nv
    r = [e % R for e in evals]
    out = []
    for i in range(nv):
        half = len(r) // 2
        p = point[i] % R
        q = [(r[2 * b + 1] - r[2 * b]) % R for b in range(half)]
        r = [(r[2 * b] * (1 - p) + r[2 * b + 1] * p) % R for b in range(half)]
        out.append(q)
    return out


def open_proofs(evals: Sequence[int], point: Sequence[int], level_bases: Sequence[Sequence], msm: Callable):
    proofs = []
    for q, bases in zip(quotients(evals, point), level_bases):
        scalars = [q[x >> 1] for x in range(2 * len(q))]
        assert len(bases) == len(scalars)
        proofs.append(msm(bases, scalars))
    return proofs


def eq_exponents(t: Sequence[int]) -> List[int]:
    """eq(t, x) for x in {0,1}^len(t), little-endian: bit j of x selects t[j] (set) or 1 - t[j] (clear)."""
    out = [1]
    for tj in t:
        out = [v * ((1 - tj) % R) % R for v in out] + [v * (tj % R) % R for v in out]
    return out


def mle_eval(evals: Sequence[int], t: Sequence[int]) -> int:
    """multilinear extension with little-endian variables: sum_x evals[x] * eq(t, x)"""
    return sum(e * w for e, w in zip(evals, eq_exponents(t))) % R
