"""CPU tests for the pairing path: the big-integer oracle pinned against the definition
(oracle/pairing.py (A) vs (B), bilinearity, non-degeneracy), the generated tower constants, and the product's
__host__ __device__ Fq12 / Miller-loop / final-exponentiation code (fq12.cuh) compiled with g++ against that oracle."""
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


def P(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def f12w(a):
    return np.array(pr.to_words(a), dtype=np.uint64)


def f2w(a):
    return np.array(o.to_limbs64(o.fq_to_mont(a[0]), 6) + o.to_limbs64(o.fq_to_mont(a[1]), 6), dtype=np.uint64)


def rand_f12(rng):
    return tuple((rng.randrange(o.Q), rng.randrange(o.Q)) for _ in range(6))


@pytest.fixture(scope="module")
def hc():
    so = os.path.join(HC_DIR, "libhostcheck.so")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-o", so, os.path.join(HC_DIR, "host_check.cpp")])
    return ctypes.CDLL(so)


@pytest.fixture(scope="module")
def e_gen():
    return pr.pairing(o.G, o2.G2)


def test_oracle_definition_equals_ark_restatement(e_gen):
    """(A) affine Miller lines over Fq12 + one big power == (B) ark's projective steps, sparse lines and 2020/875 chain."""
    assert pr.pairing_definition(o.G, o2.G2) == e_gen
    assert e_gen != pr.F12_ONE and pr.f12_pow(e_gen, o.R_ORDER) == pr.F12_ONE


def test_oracle_bilinear_and_product(e_gen):
    a, b = 0x1234567890ABCDEF1234567, 0xFEDCBA0987654321
    pa, qb = o.mul(a, o.G), o2.mul(b, o2.G2)
    assert pr.pairing(pa, qb) == pr.f12_pow(e_gen, a * b % o.R_ORDER)
    # multi_pairing is the product of the pairings; identity points contribute 1 (ark skips them)
    got = pr.multi_pairing([pa, o.G, None, o.G], [o2.G2, qb, o2.G2, None])
    assert got == pr.f12_pow(e_gen, (a + b) % o.R_ORDER)
    assert pr.multi_pairing([o.G, o.neg(o.G)], [o2.G2, o2.G2]) == pr.F12_ONE


def test_tower_constants_match_generated_header():
    """fq12_consts.inc (emitted by testudo_b200/csrc/gen_fq12_consts.py, no oracle import) holds the oracle's values."""
    import re

    txt = open(os.path.join(ROOT, "testudo_b200", "csrc", "fq12_consts.inc")).read()
    words = [int(x, 16) for x in re.findall(r"0x([0-9a-f]{8})u", txt)]
    vals = [o.fq_from_mont(sum(w << (32 * i) for i, w in enumerate(words[12 * k:12 * k + 12]))) for k in range(len(words) // 12)]
    fc = pr.frobenius_constants()
    exp = []
    for i in range(1, 6):
        exp += [fc[1][i][0], fc[1][i][1]]
    for i in range(1, 6):
        assert fc[2][i][1] == 0
        exp.append(fc[2][i][0])
    exp += [o2.B2[1], pow(2, -1, o.Q)]
    assert o2.B2[0] == 0 and vals[:-1] == exp
    # every Frobenius coefficient lies in Fq (the cooperative kernels rely on it)
    assert all(fc[1][i][1] == 0 for i in range(1, 6))
    # G1_BETA: phi(x, y) = (beta x, y) is multiplication by lambda = x^2 - 1 on G1; psi on G2 is multiplication by x
    beta, lam = vals[-1], pr.X * pr.X - 1
    pt = o.mul(0xC0FFEE123456789, o.G)
    assert (pt[0] * beta % o.Q, pt[1]) == o.mul(lam, pt) and (lam * lam + lam + 1) % o.R_ORDER == 0
    q2 = o2.mul(0xBADC0DE987654321, o2.G2)
    psi = (o2.f2_mul(pr.f2_conj(q2[0]), fc[1][2]), o2.f2_mul(pr.f2_conj(q2[1]), fc[1][3]))
    assert psi == o2.mul(pr.X, q2) and pr.X ** 4 > o.R_ORDER


def test_host_fq12_arithmetic(hc):
    rng = random.Random(7)
    for _ in range(4):
        a, b = rand_f12(rng), rand_f12(rng)
        A, B = f12w(a), f12w(b)
        out = np.zeros(72, dtype=np.uint64)
        hc.hc_fq12_mul(P(A), P(B), P(out))
        assert pr.from_words(out) == pr.f12_mul(a, b)
        hc.hc_fq12_sqr(P(A), P(out))
        assert pr.from_words(out) == pr.f12_sqr(a)
        hc.hc_fq12_inv(P(A), P(out))
        assert pr.f12_mul(pr.from_words(out), a) == pr.F12_ONE
        for k in (1, 2):
            hc.hc_fq12_frobenius(P(A), k, P(out))
            assert pr.from_words(out) == pr.f12_frobenius(a, k)
        l0, l3, l4 = rand_f12(rng)[:3]
        hc.hc_fq12_mul_by_034(P(A), P(f2w(l0)), P(f2w(l3)), P(f2w(l4)), P(out))
        assert pr.from_words(out) == pr.f12_mul(a, (l0, l3, (0, 0), l4, (0, 0), (0, 0)))
    # edge values
    one = f12w(pr.F12_ONE)
    out = np.zeros(72, dtype=np.uint64)
    hc.hc_fq12_mul(P(one), P(one), P(out))
    assert pr.from_words(out) == pr.F12_ONE


def test_host_cyclotomic_ops(hc, e_gen):
    """Granger-Scott squaring and exp_by_x are only valid on unitary elements: use GT elements."""
    g = pr.f12_pow(e_gen, 0xABCDEF123)
    G = f12w(g)
    out = np.zeros(72, dtype=np.uint64)
    hc.hc_fq12_cyclotomic_sqr(P(G), P(out))
    assert pr.from_words(out) == pr.f12_sqr(g)
    hc.hc_fq12_exp_by_x(P(G), P(out))
    assert pr.from_words(out) == pr.f12_pow(g, pr.X)


def test_host_miller_loop_and_final_exp(hc, e_gen):
    rng = random.Random(11)
    a, b = rng.randrange(1, o.R_ORDER), rng.randrange(1, o.R_ORDER)
    pa, qb = o.mul(a, o.G), o2.mul(b, o2.G2)
    Pw = np.array(o.affine_to_words(pa), dtype=np.uint64)
    Qw = np.array(o2.affine_to_words(qb), dtype=np.uint64)
    f = np.zeros(72, dtype=np.uint64)
    e = np.zeros(72, dtype=np.uint64)
    hc.hc_miller_loop(P(Pw), P(Qw), P(f))
    hc.hc_final_exp(P(f), P(e))
    exp = pr.pairing(pa, qb)
    assert pr.from_words(e) == exp == pr.f12_pow(e_gen, a * b % o.R_ORDER)
    # the Miller values themselves agree up to the subfield factors; after the oracle's final exponentiation too
    assert pr.final_exponentiation(pr.from_words(f)) == exp
    # identity operands: Miller value 1
    z1 = np.zeros(12, dtype=np.uint64)
    hc.hc_miller_loop(P(z1), P(Qw), P(f))
    assert pr.from_words(f) == pr.F12_ONE
    z2 = np.zeros(24, dtype=np.uint64)
    hc.hc_miller_loop(P(Pw), P(z2), P(f))
    assert pr.from_words(f) == pr.F12_ONE
