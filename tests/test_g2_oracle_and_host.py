"""CPU tests for the G2 path: the big-integer oracle's constants and golden vectors, and the
product's __host__ __device__ Fq2 / G2 group-law code (g2.cuh) compiled with g++ against that oracle."""
import ctypes
import json
import os
import random
import subprocess

import numpy as np
import pytest

from oracle import bls12_377 as o
from oracle import bls12_377_g2 as o2

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HC_DIR = os.path.join(ROOT, "tests", "host_check")
GOLD = json.load(open(os.path.join(ROOT, "tests", "golden", "g2_golden.json")))


def unhex(p):
    return None if p is None else ((int(p[0], 16), int(p[1], 16)), (int(p[2], 16), int(p[3], 16)))


def P(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def f2w(a):
    return np.array(o.to_limbs64(o.fq_to_mont(a[0]), 6) + o.to_limbs64(o.fq_to_mont(a[1]), 6), dtype=np.uint64)


def w2f(w):
    return (o.fq_from_mont(o.from_limbs64([int(x) for x in w[:6]])), o.fq_from_mont(o.from_limbs64([int(x) for x in w[6:]])))


def aw(p):
    return np.array(o2.affine_to_words(p), dtype=np.uint64)


@pytest.fixture(scope="module")
def hc():
    so = os.path.join(HC_DIR, "libhostcheck.so")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-o", so, os.path.join(HC_DIR, "host_check.cpp")])
    return ctypes.CDLL(so)


def test_g2_constants():
    """The curve definition is pinned numerically: -5 is a quadratic non-residue (Fq2 is a field), B' = 1/u, the
    generator is on the twist and has order r."""
    assert pow(o2.NONRESIDUE, (o.Q - 1) // 2, o.Q) == o.Q - 1
    assert o2.f2_mul(o2.B2, (0, 1)) == (1, 0)
    assert o2.is_on_curve(o2.G2)
    rm1 = o2.mul(o.R_ORDER - 1, o2.G2)
    assert rm1 == o2.neg(o2.G2) and o2.add(rm1, o2.G2) is None
    k = GOLD["curve_kats"]
    assert unhex(k["G2"]) == o2.G2 and unhex(k["2G2"]) == o2.add(o2.G2, o2.G2) and k["rG2_is_inf"] is True
    assert unhex(k["3G2"]) == o2.mul(3, o2.G2) and unhex(k["12345G2"]) == o2.mul(12345, o2.G2)
    assert unhex(k["(r-1)G2"]) == rm1


def test_g2_oracle_golden_vectors():
    for case in GOLD["explicit"] + GOLD["edge"]:
        pts = [unhex(p) for p in case["points"]]
        sc = [int(s, 16) for s in case["scalars"]]
        assert all(o2.is_on_curve(p) for p in pts)
        assert o2.msm_naive(pts, sc) == unhex(case["result"]), case.get("name")
    for case in GOLD["seeded"]:
        if case["n"] > 33:
            continue
        pts, dl = o2.rand_points(case["n"], case["points_seed"])
        sc = o.rand_scalars(case["n"], case["scalars_seed"])
        assert o2.msm_naive(pts, sc) == unhex(case["result"]) == o2.msm_by_dlog(dl, sc)
    # layout round trip
    p = o2.mul(777, o2.G2)
    assert o2.affine_from_words(o2.affine_to_words(p)) == p and o2.affine_from_words([0] * 24) is None


def test_fq2_host_arithmetic(hc):
    rng = random.Random(5)
    out = np.zeros(12, np.uint64)
    vals = [(0, 1), (1, 0), (o.Q - 1, o.Q - 1), (0, o.Q - 1), (5, 0), (0, 0)]
    vals += [(rng.randrange(o.Q), rng.randrange(o.Q)) for _ in range(300)]
    for i, a in enumerate(vals):
        b = vals[(7 * i + 3) % len(vals)]
        hc.hc_fq2_mul(P(f2w(a)), P(f2w(b)), P(out)); assert w2f(out) == o2.f2_mul(a, b)
        hc.hc_fq2_sqr(P(f2w(a)), P(out)); assert w2f(out) == o2.f2_mul(a, a)
        if a != (0, 0):
            hc.hc_fq2_inv(P(f2w(a)), P(out)); assert w2f(out) == o2.f2_inv(a)


def test_g2_host_group_law_exceptional_cases(hc):
    pts, _ = o2.rand_points(12, 3)
    T = pts[11]
    out = np.zeros(24, np.uint64)
    cases = [(pts[0], pts[1]), (pts[2], pts[2]), (pts[3], o2.neg(pts[3])), (None, pts[4]), (pts[5], None), (None, None),
             (pts[6], pts[7])]
    for p, q in cases:
        hc.hc_g2_madd(P(aw(p)), P(aw(q)), P(out)); assert o2.affine_from_words(out) == o2.add(p, q)
        hc.hc_g2_add(P(aw(p)), P(aw(q)), P(aw(T)), P(out)); assert o2.affine_from_words(out) == o2.add(p, q)
        hc.hc_g2_dbl(P(aw(p)), P(aw(T)), P(out)); assert o2.affine_from_words(out) == o2.add(p, p)
    rng = random.Random(4)
    for k in [0, 1, 2, 3, o.R_ORDER - 1, rng.randrange(o.R_ORDER)]:
        kw = np.array([(k >> (32 * i)) & 0xFFFFFFFF for i in range(8)], dtype=np.uint32)
        hc.hc_g2_scalar_mul(P(aw(pts[8])), P(kw), P(out)); assert o2.affine_from_words(out) == o2.mul(k, pts[8])
