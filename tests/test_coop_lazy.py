"""The cooperative pairing engine's lazily reduced per-item bodies (testudo_b200/csrc/fq12_coop.cuh) on the HOST: the same
functions the kernels call, run item by item in sequence (tests/host_check), against the oracle tower -- on random values,
on extreme ones (every coefficient q - 1, zero, non-canonical representatives up to 1.02 q) and through whole chains
(exp_by_x, the final exponentiation, a Miller loop). An offset that is too small or a bound that is exceeded shows up as
a wrong residue here, before any GPU time is spent."""
import ctypes
import os
import random
import subprocess

import numpy as np
import pytest

from oracle import bls12_377 as o
from oracle import bls12_377_g2 as o2
from oracle import pairing as pr

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HC_DIR = os.path.join(ROOT, "tests", "host_check")
Q = o.Q
R384 = 1 << 384
RINV = pow(R384, -1, Q)


def P(a):
    return a.ctypes.data_as(ctypes.c_void_p)


@pytest.fixture(scope="module")
def hc():
    so = os.path.join(HC_DIR, "libhostcheck.so")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-o", so, os.path.join(HC_DIR, "host_check.cpp")])
    lib = ctypes.CDLL(so)
    lib.hc_coop_op_top.restype = ctypes.c_uint32
    return lib


def raw_words(vals):
    """12 raw 384-bit integers (Montgomery representatives, not necessarily canonical) -> [72] u64"""
    out = []
    for v in vals:
        assert 0 <= v < R384
        out += [(v >> (64 * i)) & (2**64 - 1) for i in range(6)]
    return np.array(out, dtype=np.uint64)


def raw_to_f12(vals):
    """the Fq12 element (oracle order: by power of w) the 12 raw representatives (memory = tower order) stand for"""
    e = [v * RINV % Q for v in vals]
    tower = [(e[2 * i], e[2 * i + 1]) for i in range(6)]
    flat = [None] * 6
    for pos, i in enumerate((0, 2, 4, 1, 3, 5)):
        flat[i] = tower[pos]
    return tuple(flat)


def f12_to_raw(x):
    """canonical Montgomery representatives in memory order"""
    return [c * R384 % Q for i in (0, 2, 4, 1, 3, 5) for c in x[i]]


def run(hc, op, a_raw, b_raw=None):
    A = raw_words(a_raw)
    B = raw_words(b_raw if b_raw is not None else a_raw)
    out = np.zeros(72, dtype=np.uint64)
    hc.hc_coop_op(op, P(A), P(B), P(out))
    return pr.from_words(out)


def extreme_inputs(rng):
    top = Q + Q // 50 - 1                         # the invariant allows any representative < 1.02 q
    yield [Q - 1] * 12
    yield [0] * 12
    yield [top] * 12
    yield [top if i % 2 else 0 for i in range(12)]
    yield [0 if i % 2 else top for i in range(12)]
    yield [top if i < 6 else 1 for i in range(12)]
    for _ in range(6):
        yield [rng.choice([0, 1, Q - 1, Q, top, rng.randrange(Q)]) for _ in range(12)]
    for _ in range(6):
        yield [rng.randrange(top + 1) for _ in range(12)]


def test_lazy_mul_sqr_unary_on_extreme_inputs(hc):
    rng = random.Random(41)
    ins = list(extreme_inputs(rng))
    limit = (Q + Q // 50) >> 352                  # top limb of 1.02 q
    for a in ins:
        fa = raw_to_f12(a)
        assert run(hc, 1, a) == pr.f12_sqr(fa)
        assert run(hc, 3, a) == pr.f12_frobenius(fa, 1)
        assert run(hc, 4, a) == pr.f12_frobenius(fa, 2)
        assert run(hc, 5, a) == pr.f12_conj(fa)
        for op in (1, 3, 4, 5):
            assert hc.hc_coop_op_top(op, P(raw_words(a)), P(raw_words(a))) <= limit
        for b in ins[::3]:
            fb = raw_to_f12(b)
            assert run(hc, 0, a, b) == pr.f12_mul(fa, fb)
            assert hc.hc_coop_op_top(0, P(raw_words(a)), P(raw_words(b))) <= limit
            want = pr.f12_conj(pr.f12_sqr(pr.f12_mul(pr.f12_sqr(pr.f12_mul(fa, fb)), fa)))
            assert run(hc, 6, a, b) == want


def test_lazy_cyclotomic_chain_and_final_exponentiation(hc):
    e_gen = pr.pairing(o.G, o2.G2)
    rng = random.Random(43)
    limit = (Q + Q // 50) >> 352
    for k in (1, 0xABCDEF123, rng.randrange(1, o.R_ORDER)):
        g = pr.f12_pow(e_gen, k)
        raw = f12_to_raw(g)
        assert run(hc, 2, raw) == pr.f12_sqr(g)
        assert hc.hc_coop_op_top(2, P(raw_words(raw)), P(raw_words(raw))) <= limit
        assert run(hc, 7, raw) == pr.f12_pow(g, pr.X)
        # a non-canonical representative of the same unitary element
        raw2 = [c + Q if c < Q // 50 else c for c in raw]
        assert run(hc, 2, raw2) == pr.f12_sqr(g)
    for _ in range(2):
        x = tuple((rng.randrange(Q), rng.randrange(Q)) for _ in range(6))
        raw = f12_to_raw(x)
        assert run(hc, 8, raw) == pr.final_exponentiation(x)
    # degenerate shapes for the cooperative Fq6 inversion of the easy part: 1, an element of Fq6, of Fq2, a pure w multiple
    z = (0, 0)
    for x in (pr.F12_ONE, ((3, 5), z, (7, 11), z, (13, 17), z), ((Q - 1, Q - 2), z, z, z, z, z), (z, (Q - 1, 1), z, z, z, z)):
        assert run(hc, 8, f12_to_raw(x)) == pr.final_exponentiation(x)


def test_lazy_generic_power_chain(hc):
    """The chain of k_fq12_pow_coop (the verifier's `tx.pow(c)`, src/mipp.rs:258-261): ~253 generic lazily reduced squarings
    with a product after every set bit, on arbitrary field elements (proof values are not known to be unitary)."""
    rng = random.Random(47)
    cases = [(tuple((rng.randrange(Q), rng.randrange(Q)) for _ in range(6)), e)
             for e in (1, 2, 3, o.R_ORDER - 1, (1 << 253) - 1, rng.randrange(1, o.R_ORDER), 1 << 252, (1 << 256) - 1)]
    cases.append((tuple((Q - 1, Q - 1) for _ in range(6)), rng.randrange(1, o.R_ORDER)))
    for x, e in cases:
        A = raw_words(f12_to_raw(x))
        ew = np.array([(e >> (32 * i)) & 0xFFFFFFFF for i in range(8)], dtype=np.uint32)
        out = np.zeros(72, dtype=np.uint64)
        hc.hc_coop_pow(P(A), P(ew), P(out))
        assert pr.from_words(out) == pr.f12_pow(x, e)


def g2hom_words(pt):
    out = []
    for c in pt:
        for v in c:
            m = v * R384 % Q
            out += [(m >> (64 * i)) & (2**64 - 1) for i in range(6)]
    return np.array(out, dtype=np.uint64)


def test_lazy_doubling_step_equals_the_canonical_one(hc):
    rng = random.Random(47)
    cases = []
    for _ in range(8):
        cases.append(([(rng.randrange(Q), rng.randrange(Q)) for _ in range(3)], rng.randrange(Q), rng.randrange(Q)))
    cases.append(([(Q - 1, Q - 1)] * 3, Q - 1, Q - 1))
    cases.append(([(0, Q - 1), (Q - 1, 0), (1, Q - 1)], 1, Q - 1))
    cases.append(([(0, 0), (0, 0), (0, 0)], 0, 0))
    for r, px, py in cases:
        Rw = g2hom_words(r)
        pxw = raw_words([px * R384 % Q])[:6].copy()
        pyw = raw_words([py * R384 % Q])[:6].copy()
        r_out = np.zeros(36, dtype=np.uint64)
        line_out = np.zeros(36, dtype=np.uint64)
        assert hc.hc_coop_double_step(P(Rw), P(pxw), P(pyw), P(r_out), P(line_out)) == 1


def test_lazy_miller_loop_on_the_host(hc):
    rng = random.Random(53)
    a, b = rng.randrange(1, o.R_ORDER), rng.randrange(1, o.R_ORDER)
    pa, qb = o.mul(a, o.G), o2.mul(b, o2.G2)
    Pw = np.array(o.affine_to_words(pa), dtype=np.uint64)
    Qw = np.array(o2.affine_to_words(qb), dtype=np.uint64)
    f = np.zeros(72, dtype=np.uint64)
    ref = np.zeros(72, dtype=np.uint64)
    hc.hc_coop_miller(P(Pw), P(Qw), P(f))
    hc.hc_miller_loop(P(Pw), P(Qw), P(ref))
    assert np.array_equal(f, ref)                  # same line functions as the single-thread loop: the same Fq12 value
    assert pr.final_exponentiation(pr.from_words(f)) == pr.pairing(pa, qb)
    z1 = np.zeros(12, dtype=np.uint64)
    hc.hc_coop_miller(P(z1), P(Qw), P(f))
    assert pr.from_words(f) == pr.F12_ONE


def test_lazy_bodies_stress_against_the_canonical_tower(hc):
    """4000 products, squarings and doubling steps on representatives biased to 0, q - 1, q and 1.02 q - 1"""
    assert hc.hc_coop_stress(ctypes.c_uint64(0x9E3779B97F4A7C15), 4000) == 0


def test_lazy_line_product(hc):
    """(l0, 0, 0, l3, l4, 0) * (m0, 0, 0, m3, m4, 0) with 18 Fq products == the general Fq12 product"""
    rng = random.Random(59)
    cases = [[rng.randrange(Q) for _ in range(12)] for _ in range(8)] + [[Q - 1] * 12, [0] * 12, [1] * 12]
    for a in cases:
        for b in cases[::2]:
            assert hc.hc_coop_line_mul(P(raw_words(a)), P(raw_words(b))) == 1


def test_lazy_g2_xyzz_group_law(hc):
    """The lane-parallel XYZZ doubling / addition over Fq2 (window combine of a G2 MSM) against the canonical routines and
    the oracle's curve arithmetic: chains with non-trivial ZZ, equal operands, opposite operands, identities."""
    pts, _ = o2.rand_points(3, 61)
    P0, Q0, R0 = pts

    def xyzz(pt):
        if pt is None:
            return np.zeros(48, dtype=np.uint64)
        w = list(o2.affine_to_words(pt))
        one = [(R384 % Q >> (64 * i)) & (2**64 - 1) for i in range(6)]
        return np.array(w + one + [0] * 6 + one + [0] * 6, dtype=np.uint64)

    def op(code, a, b=None):
        out = np.zeros(48, dtype=np.uint64)
        assert hc.hc_coop_g2_op(code, P(a), P(b if b is not None else a), P(out)) == 1
        return out

    def to_affine(v):
        """canonical XYZZ words -> oracle affine point"""
        vals = [sum(int(v[6 * k + i]) << (64 * i) for i in range(6)) * RINV % Q for k in range(8)]
        x, y, zz, zzz = (vals[0], vals[1]), (vals[2], vals[3]), (vals[4], vals[5]), (vals[6], vals[7])
        if zz == (0, 0):
            return None
        return (o2.f2_mul(x, o2.f2_inv(zz)), o2.f2_mul(y, o2.f2_inv(zzz)))

    a = xyzz(P0)
    a = op(0, a)                                   # 2 P
    a = op(0, a)                                   # 4 P, ZZ != 1
    assert to_affine(a) == o2.mul(4, P0)
    b = op(1, a, xyzz(Q0))                         # 4 P + Q
    assert to_affine(b) == o2.add(o2.mul(4, P0), Q0)
    c = op(0, xyzz(R0))
    d = op(1, b, c)                                # both with ZZ != 1
    assert to_affine(d) == o2.add(o2.add(o2.mul(4, P0), Q0), o2.mul(2, R0))
    assert to_affine(op(1, d, d)) == o2.mul(2, to_affine(d))            # equal operands: the exact path doubles
    neg = d.copy()
    y0 = sum(int(d[12 + i]) << (64 * i) for i in range(6)); y1 = sum(int(d[18 + i]) << (64 * i) for i in range(6))
    for i in range(6):
        neg[12 + i] = ((Q - y0) % Q >> (64 * i)) & (2**64 - 1)
        neg[18 + i] = ((Q - y1) % Q >> (64 * i)) & (2**64 - 1)
    assert to_affine(op(1, d, neg)) is None                            # P + (-P)
    inf = xyzz(None)
    assert to_affine(op(1, inf, xyzz(Q0))) == Q0
    assert to_affine(op(1, d, inf)) == to_affine(d)
    assert to_affine(op(0, inf)) is None


def test_fr_pair_product_with_one_reduction(hc):
    """mont_mul2_lazy<Fr> + one conditional subtraction (the pair product of k_fr_matvec) on extreme and random operands:
    Fr has only 3 spare bits, so b + d + r < 2^256 is the bound that matters"""
    R = o.R_ORDER
    RR = 1 << 256
    rinv = pow(RR, -1, R)
    rng = random.Random(67)
    ext = [0, 1, R - 1, R - 2, 1 << 252, (1 << 252) - 1, R >> 1]

    def w(v):
        return np.array([(v >> (32 * i)) & 0xFFFFFFFF for i in range(8)], dtype=np.uint32)

    for it in range(3000):
        a, b, c, d = ([R - 1] * 4 if it == 0 else
                      [rng.choice(ext) if rng.random() < 0.4 else rng.randrange(R) for _ in range(4)])
        out = np.zeros(8, dtype=np.uint32)
        hc.hc_fr_mul2(P(w(a)), P(w(b)), P(w(c)), P(w(d)), P(out))
        assert sum(int(out[i]) << (32 * i) for i in range(8)) == (a * b + c * d) * rinv % R
