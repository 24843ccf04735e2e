"""GPU parity tests for the PST openings: tb200_pst_open_g1 / _g2 vs the Python
restatement of ark-poly-commit's `MultilinearPC::open` / `open_g1` (oracle/pst.py), and vs the closed form on a
synthetic CRS with known trapdoor at the reference's sizes."""
import ctypes

import numpy as np
import pytest

import helpers as h
from oracle import bls12_377 as o
from oracle import bls12_377_g2 as o2
from oracle import pst
from testudo_b200 import _lib, multilinear_pc

pytestmark = pytest.mark.gpu


def P(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def g2_np(points):
    return np.array([o2.affine_to_words(p) for p in points], dtype=np.uint64).reshape(-1, 24)


def crs_levels(engine, t, g2: bool):
    """powers[k][x] = eq((t_k..), x) * generator, built with the engine's scalar-multiplication test kernels."""
    levels, exps = [], []
    for k in range(len(t)):
        e = pst.eq_exponents(t[k:])
        n = len(e)
        out = np.zeros((n, 24 if g2 else 12), np.uint64)
        gen = np.tile(g2_np([o2.G2]) if g2 else h.pts_to_np([o.G]), (n, 1))
        fn = engine.tb200_test_g2_mul if g2 else engine.tb200_test_g1_mul
        _lib.check(fn(P(np.ascontiguousarray(gen)), P(h.scalars_to_np(e)), n, P(out)))
        levels.append(out)
        exps.append(e)
    return levels, exps


@pytest.mark.parametrize("nv", [1, 2, 5])
def test_open_g2_vs_restated_ark_algorithm(engine, nv):
    t = o.rand_scalars(nv, 50 + nv)
    levels, exps = crs_levels(engine, t, True)
    evals = o.rand_scalars(1 << nv, 60 + nv)
    point = o.rand_scalars(nv, 70 + nv)
    bases = [[o2.mul(e, o2.G2) for e in lv] for lv in exps]
    exp = pst.open_proofs(evals, point, bases, o2.msm_naive)
    got = multilinear_pc.open(levels, h.scalars_to_np(evals, mont=True), h.scalars_to_np(point, mont=True))
    assert [o2.affine_from_words(r) for r in got] == exp
    got_c = multilinear_pc.open(levels, h.scalars_to_np(evals), h.scalars_to_np(point), mont=False)
    assert np.array_equal(got, got_c)


@pytest.mark.parametrize("nv", [1, 3, 6])
def test_open_g1_vs_restated_ark_algorithm(engine, nv):
    t = o.rand_scalars(nv, 80 + nv)
    levels, exps = crs_levels(engine, t, False)
    evals = o.rand_scalars(1 << nv, 90 + nv)
    evals[0] = 0
    point = o.rand_scalars(nv, 100 + nv)
    bases = [[o.mul(e, o.G) for e in lv] for lv in exps]
    exp = pst.open_proofs(evals, point, bases, o.msm_naive)
    got = multilinear_pc.open_g1(levels, h.scalars_to_np(evals, mont=True), h.scalars_to_np(point, mont=True))
    assert [h.pt_from_np(r) for r in got] == exp


@pytest.mark.parametrize("g2,nv", [(True, 10), (False, 13)])
def test_open_closed_form_and_pst_identity(engine, g2, nv):
    """Reference sizes (q has m_row <= 13 variables). On the synthetic CRS proof_i = q_i(t_{i+1..}) * generator, and
    the exponents satisfy the PST identity f(t) - f(point) = sum_i (t_i - point_i) q_i(t) that `check` verifies with
    pairings."""
    t = o.rand_scalars(nv, 110 + nv)
    levels, _ = crs_levels(engine, t, g2)
    evals = o.rand_scalars(1 << nv, 120 + nv)
    point = o.rand_scalars(nv, 130 + nv)
    qs = pst.quotients(evals, point)
    dl = [pst.mle_eval(q, t[i + 1:]) for i, q in enumerate(qs)]
    assert (pst.mle_eval(evals, t) - pst.mle_eval(evals, point)) % o.R_ORDER == \
        sum((t[i] - point[i]) * dl[i] for i in range(nv)) % o.R_ORDER
    if g2:
        got = multilinear_pc.open(levels, h.scalars_to_np(evals, mont=True), h.scalars_to_np(point, mont=True))
        assert [o2.affine_from_words(r) for r in got] == [o2.mul(d, o2.G2) for d in dl]
    else:
        got = multilinear_pc.open_g1(levels, h.scalars_to_np(evals, mont=True), h.scalars_to_np(point, mont=True))
        assert [h.pt_from_np(r) for r in got] == [o.mul(d, o.G) for d in dl]


def test_open_argument_errors(engine):
    with pytest.raises(ValueError):
        multilinear_pc.open([np.zeros((2, 24), np.uint64)], np.zeros((4, 4), np.uint64), np.zeros((1, 4), np.uint64))
    out = np.zeros((1, 24), np.uint64)
    assert engine.tb200_pst_open_g2(None, 1, None, None, 0, P(out)) == -1
