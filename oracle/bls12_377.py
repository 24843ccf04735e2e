"""TEST INFRASTRUCTURE ONLY -- big-integer BLS12-377 G1 oracle (ground truth for <= 2^12 points).

This file is the CPU restatement of the arithmetic Testudo reaches through
ark-ec 0.4 `VariableBaseMSM` / ark-bls12-377 0.4 (neither crate is
vendored under /root/reference, only pinned by branch in Cargo.toml:21-24,73-80).
It restates the *published* definitions:

  * BLS12-377 G1: y^2 = x^3 + 1 over Fq, prime-order subgroup of order r,
    generator as in ark-bls12-377 `g1::Config::GENERATOR` (constants in SURVEY.md App. B).
  * MSM(bases, scalars) = sum_k scalars[k] * bases[k]; result compared as the
    canonical affine point (x, y) or the identity -- the only thing every
    Testudo call site consumes (src/sqrt_pst.rs:198 `.into_affine()`,
    src/mipp.rs:117,363, src/commitments.rs:76,85).
  * ark in-memory layout: Fq = 6 x u64 little-endian limbs in Montgomery form
    (R = 2^384), Fr = 4 x u64 (R = 2^256); identity carried as a flag.

PARITY UNPINNED: the reference holds no golden vectors / known-answer tests for
any MSM and cannot be compiled here (no Rust). This
oracle is pinned by curve KATs (on-curve generator, r*G = inf, (r-1)*G = -G,
2G/3G from the affine formulas) and by the uniqueness of the group law: an MSM
result is a single well-defined group element, so any correct implementation
produces the same affine coordinates as arkworks.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import
this module. The product (testudo_b200/) never does.
"""
from __future__ import annotations

import random
from typing import Iterable, List, Optional, Sequence, Tuple

# --- constants (re-verified by tests/test_oracle.py) -------------------
Q = 0x01AE3A4617C510EAC63B05C06CA1493B1A22D9F300F5138F1EF3622FBA094800170B5D44300000008508C00000000001
R_ORDER = 0x12AB655E9A2CA55660B44D1E5C37B00159AA76FED00000010A11800000000001
GX = 0x008848DEFE740A67C8FC6225BF87FF5485951E2CAA9D41BB188282C8BD37CB5CD5481512FFCD394EEAB9B16EB21BE9EF
GY = 0x01914A69C5102EFF1F674F5D30AFEEC4BD7FB348CA3E52D96D182AD44FB82305C2FE3D3634A9591AFD82DE55559C8EA6
COEFF_B = 1
FQ_LIMBS64 = 6
FR_LIMBS64 = 4
FQ_R = (1 << 384) % Q  # Montgomery radix for Fq (ark-ff Fp384<MontBackend>)
FR_R = (1 << 256) % R_ORDER  # Montgomery radix for Fr (Fp256)
FQ_RINV = pow(FQ_R, -1, Q)
FR_RINV = pow(FR_R, -1, R_ORDER)
SCALAR_BITS = 253  # Fr::MODULUS_BIT_SIZE

Affine = Optional[Tuple[int, int]]  # None == point at infinity
G: Affine = (GX, GY)


# --- affine group law -------------------------------------------------------------------------
def is_on_curve(p: Affine) -> bool:
    if p is None:
        return True
    x, y = p
    return (y * y - x * x * x - COEFF_B) % Q == 0


def neg(p: Affine) -> Affine:
    if p is None:
        return None
    return (p[0], (-p[1]) % Q)


def add(p: Affine, q: Affine) -> Affine:
    if p is None:
        return q
    if q is None:
        return p
    x1, y1 = p
    x2, y2 = q
    if x1 == x2:
        if (y1 + y2) % Q == 0:
            return None
        lam = (3 * x1 * x1) * pow(2 * y1, -1, Q) % Q
    else:
        lam = (y2 - y1) * pow(x2 - x1, -1, Q) % Q
    x3 = (lam * lam - x1 - x2) % Q
    y3 = (lam * (x1 - x3) - y1) % Q
    return (x3, y3)


# --- Jacobian arithmetic for speed (a = 0) ----------------------------------------------------
def _jdbl(p):
    X, Y, Z = p
    if Z == 0:
        return p
    A = X * X % Q
    B = Y * Y % Q
    C = B * B % Q
    D = 2 * ((X + B) * (X + B) - A - C) % Q
    E = 3 * A % Q
    F = E * E % Q
    X3 = (F - 2 * D) % Q
    Y3 = (E * (D - X3) - 8 * C) % Q
    Z3 = 2 * Y * Z % Q
    return (X3, Y3, Z3)


def _jadd(p, q):
    X1, Y1, Z1 = p
    X2, Y2, Z2 = q
    if Z1 == 0:
        return q
    if Z2 == 0:
        return p
    Z1Z1 = Z1 * Z1 % Q
    Z2Z2 = Z2 * Z2 % Q
    U1 = X1 * Z2Z2 % Q
    U2 = X2 * Z1Z1 % Q
    S1 = Y1 * Z2 * Z2Z2 % Q
    S2 = Y2 * Z1 * Z1Z1 % Q
    if U1 == U2:
        if S1 == S2:
            return _jdbl(p)
        return (1, 1, 0)
    H = (U2 - U1) % Q
    Rr = (S2 - S1) % Q
    HH = H * H % Q
    HHH = H * HH % Q
    V = U1 * HH % Q
    X3 = (Rr * Rr - HHH - 2 * V) % Q
    Y3 = (Rr * (V - X3) - S1 * HHH) % Q
    Z3 = Z1 * Z2 * H % Q
    return (X3, Y3, Z3)


def _to_jac(p: Affine):
    return (1, 1, 0) if p is None else (p[0], p[1], 1)


def _from_jac(p) -> Affine:
    X, Y, Z = p
    if Z == 0:
        return None
    zi = pow(Z, -1, Q)
    zi2 = zi * zi % Q
    return (X * zi2 % Q, Y * zi2 * zi % Q)


def mul(k: int, p: Affine) -> Affine:
    """k * p for any integer k (reduced mod r only if p is in the subgroup; we reduce sign only)."""
    if p is None or k == 0:
        return None
    if k < 0:
        return mul(-k, neg(p))
    acc = (1, 1, 0)
    base = _to_jac(p)
    for bit in bin(k)[2:]:
        acc = _jdbl(acc)
        if bit == "1":
            acc = _jadd(acc, base)
    return _from_jac(acc)


def msm_naive(bases: Sequence[Affine], scalars: Sequence[int]) -> Affine:
    """Definition of the MSM; `msm_unchecked` semantics: truncates to the shorter input
    (ark-ec 0.4 `VariableBaseMSM::msm_unchecked`)."""
    n = min(len(bases), len(scalars))
    acc = (1, 1, 0)
    for i in range(n):
        s = scalars[i] % R_ORDER
        if s == 0 or bases[i] is None:
            continue
        acc = _jadd(acc, _to_jac(mul(s, bases[i])))
    return _from_jac(acc)


def msm_checked(bases: Sequence[Affine], scalars: Sequence[int]):
    """`VariableBaseMSM::msm`: Ok(point) iff lengths match else Err(min_len) (App. A.1)."""
    if len(bases) != len(scalars):
        return ("err", min(len(bases), len(scalars)))
    return ("ok", msm_naive(bases, scalars))


# --- ark memory layout helpers ------------------------------------------------------------------
def to_limbs64(v: int, n: int) -> List[int]:
    return [(v >> (64 * i)) & 0xFFFFFFFFFFFFFFFF for i in range(n)]


def from_limbs64(limbs: Iterable[int]) -> int:
    out = 0
    for i, l in enumerate(limbs):
        out |= int(l) << (64 * i)
    return out


def fq_to_mont(v: int) -> int:
    return v * FQ_R % Q


def fq_from_mont(v: int) -> int:
    return v * FQ_RINV % Q


def fr_to_mont(v: int) -> int:
    return v * FR_R % R_ORDER


def fr_from_mont(v: int) -> int:
    return v * FR_RINV % R_ORDER


def affine_to_words(p: Affine) -> List[int]:
    """96-byte C-ABI encoding: x[6] || y[6] u64 LE limbs, Montgomery form; identity == all zero
    ((0,0) is not on y^2 = x^3 + 1, so the encoding is unambiguous)."""
    if p is None:
        return [0] * 12
    return to_limbs64(fq_to_mont(p[0]), 6) + to_limbs64(fq_to_mont(p[1]), 6)


def affine_from_words(w: Sequence[int]) -> Affine:
    xm = from_limbs64(w[0:6])
    ym = from_limbs64(w[6:12])
    if xm == 0 and ym == 0:
        return None
    return (fq_from_mont(xm), fq_from_mont(ym))


# --- deterministic synthetic inputs -------------------------------------------------
def rand_scalars(n: int, seed: int) -> List[int]:
    rng = random.Random(seed)
    return [rng.randrange(R_ORDER) for _ in range(n)]


def rand_points(n: int, seed: int) -> Tuple[List[Affine], List[int]]:
    """n subgroup points with known discrete logs: P_k = d_k * G, built incrementally so the
    cost is one add per point: d_k = d_0 + k*step."""
    rng = random.Random(seed ^ 0x5125_0001)
    d0 = rng.randrange(1, R_ORDER)
    step = rng.randrange(1, R_ORDER)
    p = _to_jac(mul(d0, G))
    sp = _to_jac(mul(step, G))
    pts: List[Affine] = []
    dl: List[int] = []
    d = d0
    jac = []
    for _ in range(n):
        jac.append(p)
        dl.append(d)
        p = _jadd(p, sp)
        d = (d + step) % R_ORDER
    # batch-normalise with Montgomery's trick
    zs = [j[2] for j in jac]
    pref = [1]
    for z in zs:
        pref.append(pref[-1] * (z if z else 1) % Q)
    inv = pow(pref[-1], -1, Q)
    for i in range(n - 1, -1, -1):
        z = zs[i]
        if z == 0:
            pts.append(None)
            continue
        zi = inv * pref[i] % Q
        inv = inv * z % Q
        zi2 = zi * zi % Q
        pts.append((jac[i][0] * zi2 % Q, jac[i][1] * zi2 * zi % Q))
    pts.reverse()
    return pts, dl


def msm_by_dlog(dlogs: Sequence[int], scalars: Sequence[int]) -> Affine:
    """Closed-form MSM for bases with known discrete logs: (sum s_k d_k mod r) * G."""
    n = min(len(dlogs), len(scalars))
    t = 0
    for i in range(n):
        t += scalars[i] * dlogs[i]
    return mul(t % R_ORDER, G)


# --- signed-digit recoding restated from ark-ec 0.4 `make_digits` (App. A.1) ------------------------
def ark_window_bits(n: int) -> int:
    """ark-ec 0.4 msm_bigint_wnaf: c = 3 if n < 32 else ln_without_floats(n) + 2."""
    if n < 32:
        return 3
    # ark_std::log2(x) = ceil(log2(x)) (0 for x<=1); ln_without_floats(a) = log2(a) * 69 / 100
    lg = (n - 1).bit_length()
    return lg * 69 // 100 + 2


def ark_make_digits(s: int, w: int, num_bits: int = SCALAR_BITS) -> List[int]:
    radix = 1 << w
    window_mask = radix - 1
    digits_count = (num_bits + w - 1) // w
    carry = 0
    out = []
    for i in range(digits_count):
        coef = carry + ((s >> (w * i)) & window_mask)
        carry = (coef + radix // 2) >> w
        d = coef - (carry << w)
        if i == digits_count - 1:
            d += carry << w
        out.append(d)
    return out


def msm_pippenger(bases: Sequence[Affine], scalars: Sequence[int], c: Optional[int] = None) -> Affine:
    """Pure-Python restatement of ark-ec 0.4 `msm_bigint_wnaf` (App. A.1) -- small cases only."""
    n = min(len(bases), len(scalars))
    if c is None:
        c = ark_window_bits(n)
    digs = [ark_make_digits(scalars[i] % R_ORDER, c) for i in range(n)]
    nwin = (SCALAR_BITS + c - 1) // c
    window_sums = []
    for w in range(nwin):
        buckets = [(1, 1, 0)] * (1 << c)
        for i in range(n):
            d = digs[i][w]
            if bases[i] is None or d == 0:
                continue
            if d > 0:
                buckets[d - 1] = _jadd(buckets[d - 1], _to_jac(bases[i]))
            else:
                buckets[-d - 1] = _jadd(buckets[-d - 1], _to_jac(neg(bases[i])))
        running = (1, 1, 0)
        res = (1, 1, 0)
        for b in reversed(buckets):
            running = _jadd(running, b)
            res = _jadd(res, running)
        window_sums.append(res)
    total = (1, 1, 0)
    for w in reversed(range(1, nwin)):
        total = _jadd(total, window_sums[w])
        for _ in range(c):
            total = _jdbl(total)
    total = _jadd(total, window_sums[0])
    return _from_jac(total)
