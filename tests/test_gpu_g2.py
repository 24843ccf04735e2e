"""GPU parity tests for the G2 MSM path: C ABI -> CUDA kernels vs the big-integer oracle
(oracle/bls12_377_g2.py) and the committed golden vectors; bit-exact (identical affine limbs)."""
import ctypes
import json
import os

import numpy as np
import pytest

import helpers as h
from oracle import bls12_377 as o
from oracle import bls12_377_g2 as o2
from testudo_b200 import _lib, msm_g2

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = json.load(open(os.path.join(ROOT, "tests", "golden", "g2_golden.json")))


def unhex(p):
    return None if p is None else ((int(p[0], 16), int(p[1], 16)), (int(p[2], 16), int(p[3], 16)))


def P(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def pts_np(points):
    return np.array([o2.affine_to_words(p) for p in points], dtype=np.uint64).reshape(-1, 24)


def test_g2_group_kernels(engine):
    pts, _ = o2.rand_points(40, 11)
    a = pts[:20] + [None, pts[3], pts[5], None]
    b = pts[20:] + [pts[1], pts[3], o2.neg(pts[5]), None]
    out = np.zeros((len(a), 24), np.uint64)
    _lib.check(engine.tb200_test_g2_add(P(pts_np(a)), P(pts_np(b)), len(a), P(out)))
    assert [o2.affine_from_words(r) for r in out] == [o2.add(x, y) for x, y in zip(a, b)]
    ks = [0, 1, 2, o.R_ORDER - 1] + o.rand_scalars(12, 12)
    out = np.zeros((len(ks), 24), np.uint64)
    _lib.check(engine.tb200_test_g2_mul(P(pts_np(pts[:len(ks)])), P(h.scalars_to_np(ks)), len(ks), P(out)))
    assert [o2.affine_from_words(r) for r in out] == [o2.mul(k, p) for k, p in zip(ks, pts)]


@pytest.mark.parametrize("case", GOLD["seeded"], ids=lambda c: f"n{c['n']}")
def test_g2_msm_seeded_golden(engine, case):
    pts, _ = o2.rand_points(case["n"], case["points_seed"])
    sc = o.rand_scalars(case["n"], case["scalars_seed"])
    exp = unhex(case["result"])
    assert o2.affine_from_words(msm_g2.msm_bigint(pts_np(pts), h.scalars_to_np(sc))) == exp
    assert o2.affine_from_words(msm_g2.msm_unchecked(pts_np(pts), h.scalars_to_np(sc, mont=True))) == exp


@pytest.mark.parametrize("case", GOLD["explicit"] + GOLD["edge"], ids=lambda c: c.get("name", "explicit"))
def test_g2_msm_explicit_and_edge_golden(engine, case):
    pts = [unhex(p) for p in case["points"]]
    sc = [int(s, 16) for s in case["scalars"]]
    assert o2.affine_from_words(msm_g2.msm_bigint(pts_np(pts), h.scalars_to_np(sc))) == unhex(case["result"])


def test_g2_msm_length_rules_and_empty(engine):
    pts, _ = o2.rand_points(4, 1)
    B = pts_np(pts)
    S = h.scalars_to_np([1, 2, 3], mont=True)
    assert msm_g2.msm(B, S) == ("err", 3)
    assert o2.affine_from_words(msm_g2.msm_unchecked(B, S)) == o2.msm_naive(pts[:3], [1, 2, 3])
    assert o2.affine_from_words(msm_g2.msm_bigint(B[:0], S[:0])) is None


@pytest.mark.parametrize("c", [0, 4, 9, 13])
def test_g2_msm_closed_form_2p13(engine, c):
    """The reference's largest G2 MSM shape (2^13 points, BASELINE configs[2]) against the closed form, for the
    automatic and for forced window widths; skewed scalars exercise multi-segment buckets."""
    n = 1 << 13
    pts, dl = o2.rand_points(n, 21)
    rng = np.random.default_rng(5)
    sc = o.rand_scalars(n, 22)
    for i in range(0, n, 2):
        sc[i] = int(rng.integers(0, 2))            # 0 / 1 heavy, like R1CS witnesses
    engine.tb200_set_window_bits(c)
    try:
        got = msm_g2.msm_bigint(pts_np(pts), h.scalars_to_np(sc))
    finally:
        engine.tb200_set_window_bits(0)
    assert o2.affine_from_words(got) == o2.msm_by_dlog(dl, sc)


def test_g2_compress(engine):
    pts, _ = o2.rand_points(16, 31)
    c = o.rand_scalars(1, 32)[0]
    got = msm_g2.compress(pts_np(pts), 8, h.scalars_to_np([c], mont=True)[0], mont=True)
    assert [o2.affine_from_words(r) for r in got] == [o2.add(pts[i], o2.mul(c, pts[8 + i])) for i in range(8)]
    got = msm_g2.compress(pts_np(pts), 8, h.scalars_to_np([c])[0], mont=False)
    assert [o2.affine_from_words(r) for r in got] == [o2.add(pts[i], o2.mul(c, pts[8 + i])) for i in range(8)]
