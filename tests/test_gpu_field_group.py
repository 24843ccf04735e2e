"""GPU parity: device field / group arithmetic (through the C ABI's kernel test hooks) vs the oracle."""
import ctypes
import random

import numpy as np
import pytest

import helpers as h
from oracle import bls12_377 as o
from testudo_b200 import _lib

pytestmark = pytest.mark.gpu


def P(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def test_fq_mul_add_sub_bit_exact(engine, oracle_c):
    rng = random.Random(11)
    edge = [0, 1, o.Q - 1, o.Q - 2, o.FQ_R, 1 << 376, 0xFFFFFFFF, 1 << 32, (1 << 64) - 1, o.Q >> 1, (1 << 377) - 1 - (1 << 376)]
    n = 20000
    a = edge + [rng.randrange(o.Q) for _ in range(n - len(edge))]
    b = [a[(i * 7 + 3) % n] for i in range(n)]
    A, B = h.fq_to_np(a), h.fq_to_np(b)
    out = np.zeros_like(A); oadd = np.zeros_like(A); osub = np.zeros_like(A)
    _lib.check(engine.tb200_test_fq_mul(P(A), P(B), n, P(out)))
    _lib.check(engine.tb200_test_fq_addsub(P(A), P(B), n, P(oadd), P(osub)))
    got = h.fq_from_np(out); ga = h.fq_from_np(oadd); gs = h.fq_from_np(osub)
    for i in range(n):
        assert got[i] == a[i] * b[i] * o.FQ_RINV % o.Q, i
        assert ga[i] == (a[i] + b[i]) % o.Q
        assert gs[i] == (a[i] - b[i]) % o.Q
    # and against the C oracle's 64-bit-limb CIOS on the same words
    for i in range(0, n, 997):
        assert np.array_equal(oracle_c.fq_binop("mul", A[i], B[i]), out[i])


def test_g1_madd_including_exceptional_cases(engine):
    pts, _ = o.rand_points(300, 21)
    ps = pts[:256]
    qs = pts[1:257]
    ps += [pts[5], pts[6], None, pts[8], None]
    qs += [pts[5], o.neg(pts[6]), pts[7], None, None]   # P+P, P+(-P), inf+P, P+inf, inf+inf
    n = len(ps)
    out = np.zeros((n, 12), np.uint64)
    _lib.check(engine.tb200_test_g1_add(P(h.pts_to_np(ps)), P(h.pts_to_np(qs)), n, P(out)))
    for i in range(n):
        assert h.pt_from_np(out[i]) == o.add(ps[i], qs[i]), i


def test_g1_scalar_mul(engine, oracle_c):
    pts, _ = o.rand_points(64, 22)
    ks = [0, 1, 2, o.R_ORDER - 1, 1 << 252] + o.rand_scalars(59, 23)
    Pn, Kn = h.pts_to_np(pts), h.scalars_to_np(ks)
    out = np.zeros((64, 12), np.uint64)
    _lib.check(engine.tb200_test_g1_mul(P(Pn), P(Kn), 64, P(out)))
    for i in range(64):
        assert np.array_equal(out[i], oracle_c.g1_mul(Pn[i], Kn[i])), i
    assert h.pt_from_np(out[3]) == o.neg(pts[3])
