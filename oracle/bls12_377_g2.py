"""TEST INFRASTRUCTURE ONLY (never imported by testudo_b200/): big-integer ground truth for BLS12-377 G2.

Restates, from the published curve definition, what ark-bls12-377 0.4 `g2::Config` / `Fq2Config` (un-vendored
dependency of the reference, Cargo.toml:24) compute under the reference's G2 multi-scalar multiplications:
`MultilinearPC::open` (src/sqrt_pst.rs:225), `commit_g2` / G2 `compress` of MIPP (src/mipp.rs:114,133).

    Fq2 = Fq[u] / (u^2 + 5)                        (NONRESIDUE = -5)
    E'(Fq2): y^2 = x^3 + B',  B' = (0, 1551986...874906) = 1/u   (D-type sextic twist of y^2 = x^3 + 1)
    generator: ark's G2_GENERATOR_{X,Y}_{C0,C1}

Parity is unpinned against the arkworks binary (no Rust toolchain, no golden vectors in the reference: DESIGN.md 2);
the constants are pinned numerically instead: -5 is a non-residue, B' == 1/u, the generator satisfies the curve
equation and r * G2 == identity (tests/test_oracle.py::test_g2_constants). Elements are pairs (c0, c1) of ints;
points are None (identity) or ((x0, x1), (y0, y1)). Affine chord-and-tangent arithmetic = the definition.
"""
from __future__ import annotations

import random
from typing import List, Optional, Sequence, Tuple

from . import bls12_377 as g1

Q = g1.Q
R_ORDER = g1.R_ORDER
NONRESIDUE = (-5) % Q
B2 = (0, 155198655607781456406391640216936120121836107652948796323930557600032281009004493664981332883744016074664192874906)
G2X = (233578398248691099356572568220835526895379068987715365179118596935057653620464273615301663571204657964920925606294,
       140913150380207355837477652521042157274541796891053068589147167627541651775299824604154852141315666357241556069118)
G2Y = (63160294768292073209381361943935198908131692476676907196754037919244929611450776219210369229519898517858833747423,
       149157405641012693445398062341192467754805999074082136895788947234480009303640899064710353187729182149407503257491)
G2 = (G2X, G2Y)

F2 = Tuple[int, int]
Affine2 = Optional[Tuple[F2, F2]]


def f2_add(a: F2, b: F2) -> F2:
    return ((a[0] + b[0]) % Q, (a[1] + b[1]) % Q)


def f2_sub(a: F2, b: F2) -> F2:
    return ((a[0] - b[0]) % Q, (a[1] - b[1]) % Q)


def f2_neg(a: F2) -> F2:
    return ((-a[0]) % Q, (-a[1]) % Q)


def f2_mul(a: F2, b: F2) -> F2:
    return ((a[0] * b[0] + NONRESIDUE * a[1] * b[1]) % Q, (a[0] * b[1] + a[1] * b[0]) % Q)


def f2_inv(a: F2) -> F2:
    n = (a[0] * a[0] - NONRESIDUE * a[1] * a[1]) % Q
    ni = pow(n, -1, Q)
    return (a[0] * ni % Q, (-a[1]) * ni % Q)


def is_on_curve(p: Affine2) -> bool:
    if p is None:
        return True
    x, y = p
    return f2_mul(y, y) == f2_add(f2_mul(f2_mul(x, x), x), B2)


def neg(p: Affine2) -> Affine2:
    return None if p is None else (p[0], f2_neg(p[1]))


def add(p: Affine2, q: Affine2) -> Affine2:
    if p is None:
        return q
    if q is None:
        return p
    if p[0] == q[0]:
        if f2_add(p[1], q[1]) == (0, 0):
            return None
        lam = f2_mul(f2_mul((3, 0), f2_mul(p[0], p[0])), f2_inv(f2_mul((2, 0), p[1])))
    else:
        lam = f2_mul(f2_sub(q[1], p[1]), f2_inv(f2_sub(q[0], p[0])))
    x = f2_sub(f2_sub(f2_mul(lam, lam), p[0]), q[0])
    y = f2_sub(f2_mul(lam, f2_sub(p[0], x)), p[1])
    return (x, y)


def mul(k: int, p: Affine2) -> Affine2:
    k %= R_ORDER
    r = None
    while k:
        if k & 1:
            r = add(r, p)
        p = add(p, p)
        k >>= 1
    return r


def msm_naive(bases: Sequence[Affine2], scalars: Sequence[int]) -> Affine2:
    acc = None
    for b, s in zip(bases, scalars):
        acc = add(acc, mul(s, b))
    return acc


def rand_points(n: int, seed: int) -> Tuple[List[Affine2], List[int]]:
    """n subgroup points with known discrete logs: P_i = (a + i * step) * G2 by repeated addition."""
    rng = random.Random(seed)
    a = rng.randrange(1, R_ORDER)
    step = rng.randrange(1, R_ORDER)
    p = mul(a, G2)
    s = mul(step, G2)
    pts, dl = [], []
    for i in range(n):
        pts.append(p)
        dl.append((a + i * step) % R_ORDER)
        p = add(p, s)
    return pts, dl


def msm_by_dlog(dlogs: Sequence[int], scalars: Sequence[int]) -> Affine2:
    return mul(sum(d * s for d, s in zip(dlogs, scalars)) % R_ORDER, G2)


def affine_to_words(p: Affine2) -> List[int]:
    """ark in-memory layout: x.c0 || x.c1 || y.c0 || y.c1, 6 little-endian u64 limbs each, Montgomery form;
    identity = 24 zero words (the C ABI's convention; ark keeps a separate `infinity` flag)."""
    if p is None:
        return [0] * 24
    out: List[int] = []
    for c in (p[0][0], p[0][1], p[1][0], p[1][1]):
        out += g1.to_limbs64(g1.fq_to_mont(c), 6)
    return out


def affine_from_words(w: Sequence[int]) -> Affine2:
    w = [int(x) for x in w]
    if not any(w):
        return None
    c = [g1.fq_from_mont(g1.from_limbs64(w[6 * i:6 * i + 6])) for i in range(4)]
    return ((c[0], c[1]), (c[2], c[3]))
