"""GPU parity tests for the pairing products: C ABI -> CUDA kernels vs the big-integer oracle
(oracle/pairing.py), the committed golden vectors and closed forms e(G1, G2)^(sum a_i b_i); bit-exact (identical Fq12
limbs in ark's in-memory order)."""
import ctypes
import json
import os

import numpy as np
import pytest

import helpers as h
from oracle import bls12_377 as o
from oracle import bls12_377_g2 as o2
from oracle import pairing as pr
from testudo_b200 import _lib, pairing

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = json.load(open(os.path.join(ROOT, "tests", "golden", "pairing_golden.json")))


def gt_unhex(hx):
    v = [int(x, 16) for x in hx]
    return tuple((v[2 * i], v[2 * i + 1]) for i in range(6))


def g2_np(points):
    return np.array([o2.affine_to_words(p) for p in points], dtype=np.uint64).reshape(-1, 24)


def P(a):
    return a.ctypes.data_as(ctypes.c_void_p)


E_GEN = gt_unhex(GOLD["e_G1_G2"])


def f12_np(vals):
    return np.array([pr.to_words(v) for v in vals], dtype=np.uint64).reshape(-1, 72)


def run_op(engine, op, a, b=None):
    b = a if b is None else b
    out = np.zeros_like(a)
    _lib.check(engine.tb200_test_fq12_op(op, P(a), P(b), len(a), P(out)))
    return [pr.from_words(r) for r in out]


def test_fq12_ops_on_device(engine):
    """Every tower operation of fq12.cuh, one per thread, against the oracle (op codes: kernels_pairing.cuh)."""
    import random

    rng = random.Random(3)
    n = 40
    xs = [tuple((rng.randrange(o.Q), rng.randrange(o.Q)) for _ in range(6)) for _ in range(n)]
    ys = [tuple((rng.randrange(o.Q), rng.randrange(o.Q)) for _ in range(6)) for _ in range(n)]
    xs[0], ys[1] = pr.F12_ONE, pr.F12_ONE
    A, B = f12_np(xs), f12_np(ys)
    assert run_op(engine, 0, A, B) == [pr.f12_mul(x, y) for x, y in zip(xs, ys)]
    assert run_op(engine, 1, A) == [pr.f12_sqr(x) for x in xs]
    assert all(pr.f12_mul(i, x) == pr.F12_ONE for i, x in zip(run_op(engine, 2, A), xs))
    assert run_op(engine, 3, A) == [pr.f12_frobenius(x, 1) for x in xs]
    assert run_op(engine, 4, A) == [pr.f12_frobenius(x, 2) for x in xs]
    assert run_op(engine, 8, A, B) == [pr.f12_mul(x, (y[0], y[2], (0, 0), y[4], (0, 0), (0, 0))) for x, y in zip(xs, ys)]
    # unitary elements for the cyclotomic operations
    us = [pr.f12_pow(E_GEN, rng.randrange(1, o.R_ORDER)) for _ in range(4)]
    U = f12_np(us)
    assert run_op(engine, 5, U) == [pr.f12_sqr(u) for u in us]
    assert run_op(engine, 6, U) == [pr.f12_pow(u, pr.X) for u in us]
    assert run_op(engine, 7, A[:4]) == [pr.final_exponentiation(x) for x in xs[:4]]


def test_warp_cooperative_fq12_ops(engine):
    """The warp-cooperative engine behind k_final_exp (one warp per Fq12 operation, operands in shared memory)."""
    import random

    rng = random.Random(5)
    n = 6
    xs = [tuple((rng.randrange(o.Q), rng.randrange(o.Q)) for _ in range(6)) for _ in range(n)]
    ys = [tuple((rng.randrange(o.Q), rng.randrange(o.Q)) for _ in range(6)) for _ in range(n)]
    xs[0] = pr.F12_ONE
    A, B = f12_np(xs), f12_np(ys)
    assert run_op(engine, 20, A, B) == [pr.f12_mul(x, y) for x, y in zip(xs, ys)]
    assert run_op(engine, 21, A) == [pr.f12_sqr(x) for x in xs]
    assert run_op(engine, 24, A) == [pr.f12_frobenius(x, 1) for x in xs]
    assert run_op(engine, 25, A) == [pr.f12_frobenius(x, 2) for x in xs]
    assert run_op(engine, 26, A) == [pr.f12_sqr(x) for x in xs]
    us = [pr.f12_pow(E_GEN, rng.randrange(1, o.R_ORDER)) for _ in range(3)]
    U = f12_np(us)
    assert run_op(engine, 22, U) == [pr.f12_sqr(u) for u in us]
    assert run_op(engine, 27, U) == [pr.f12_conj(pr.f12_sqr(u)) for u in us]
    assert run_op(engine, 23, A[:3]) == [pr.final_exponentiation(x) for x in xs[:3]]


def test_cooperative_miller_loop_and_both_launch_modes(engine):
    """Warp-per-pair Miller loop (op 28) against the oracle, and tb200_multi_pairing under both launch modes
    (thread-per-pair / warp-per-pair) on the same inputs, identities included."""
    ps, _ = o.rand_points(5, 19)
    qs, _ = o2.rand_points(5, 20)
    ps[3], qs[4] = None, None
    a = np.zeros((5, 72), dtype=np.uint64)
    a[:, :12] = h.pts_to_np(ps)
    a[:, 12:36] = g2_np(qs)
    got = run_op(engine, 28, a)
    assert [pr.final_exponentiation(f) for f in got] == [pr.pairing(p, q) for p, q in zip(ps, qs)]
    assert got[3] == pr.F12_ONE and got[4] == pr.F12_ONE
    ps, _ = o.rand_points(33, 21)
    qs, _ = o2.rand_points(33, 22)
    ps[7], qs[9] = None, None
    A, B = h.pts_to_np(ps), g2_np(qs)
    try:
        engine.tb200_set_pairing_coop_max(0)
        thread_mode = pairing.multi_pairing(A, B)
        engine.tb200_set_pairing_coop_max(1 << 20)
        warp_mode = pairing.multi_pairing(A, B)
    finally:
        engine.tb200_set_pairing_coop_max(8192)
    assert np.array_equal(thread_mode, warp_mode)
    assert pr.from_words(warp_mode) == pr.multi_pairing(ps, qs)


def test_miller_loop_on_device(engine):
    ps, _ = o.rand_points(3, 9)
    qs, _ = o2.rand_points(3, 10)
    a = np.zeros((3, 72), dtype=np.uint64)
    a[:, :12] = h.pts_to_np(ps)
    a[:, 12:36] = g2_np(qs)
    got = run_op(engine, 9, a)
    assert [pr.final_exponentiation(f) for f in got] == [pr.pairing(p, q) for p, q in zip(ps, qs)]


def test_pairing_of_generators(engine):
    got = pairing.pairing(h.pts_to_np([o.G])[0], g2_np([o2.G2])[0])
    assert pr.from_words(got) == E_GEN


@pytest.mark.parametrize("case", GOLD["seeded"], ids=lambda c: f"n{c['n']}")
def test_multi_pairing_seeded_golden(engine, case):
    ps, _ = o.rand_points(case["n"], case["g1_seed"])
    qs, _ = o2.rand_points(case["n"], case["g2_seed"])
    got = pairing.multi_pairing(h.pts_to_np(ps), g2_np(qs))
    assert pr.from_words(got) == gt_unhex(case["result"])


def test_multi_pairing_edge_cases(engine):
    ps, _ = o.rand_points(4, 5)
    qs, _ = o2.rand_points(4, 6)
    one = pr.F12_ONE
    # empty product, identities on either side (ark skips those pairs), cancelling pairs
    assert pr.from_words(pairing.multi_pairing(np.zeros((0, 12), np.uint64), np.zeros((0, 24), np.uint64))) == one
    assert pr.from_words(pairing.multi_pairing(h.pts_to_np([None, ps[1]]), g2_np([qs[0], None]))) == one
    got = pairing.multi_pairing(h.pts_to_np([ps[0], None, ps[2]]), g2_np([qs[0], qs[1], None]))
    assert pr.from_words(got) == pr.pairing(ps[0], qs[0])
    got = pairing.multi_pairing(h.pts_to_np([ps[0], o.neg(ps[0])]), g2_np([qs[0], qs[0]]))
    assert pr.from_words(got) == one
    got = pairing.multi_pairing(h.pts_to_np([ps[0], ps[0]]), g2_np([qs[0], o2.neg(qs[0])]))
    assert pr.from_words(got) == one
    # zip semantics: the shorter side bounds the product
    got = pairing.multi_pairing(h.pts_to_np(ps[:3]), g2_np(qs[:2]))
    assert pr.from_words(got) == pr.multi_pairing(ps[:2], qs[:2])


@pytest.mark.parametrize("n", [257, 4096])
def test_multi_pairing_closed_form_large(engine, n):
    """Beyond what the oracle finishes in seconds: points with known discrete logs, result e(G1,G2)^(sum a_i b_i)."""
    ps, dl1 = o.rand_points(n, 31)
    qs, dl2 = o2.rand_points(n, 32)
    got = pairing.multi_pairing(h.pts_to_np(ps), g2_np(qs))
    assert pr.from_words(got) == pr.f12_pow(E_GEN, sum(a * b for a, b in zip(dl1, dl2)) % o.R_ORDER)


def test_bilinearity_on_device(engine):
    a, b = 0x1F2E3D4C5B6A79880123456789, 0x0FEDCBA9876543211234
    lhs = pairing.pairing(h.pts_to_np([o.mul(a, o.G)])[0], g2_np([o2.mul(b, o2.G2)])[0])
    rhs = pairing.gt_pow(np.array(pr.to_words(E_GEN), dtype=np.uint64), h.scalars_to_np([a * b % o.R_ORDER]))[0]
    assert np.array_equal(lhs, rhs)


def test_gt_pow(engine):
    g = pr.f12_pow(E_GEN, 0xC0FFEE)
    exps = [0, 1, 2, o.R_ORDER - 1, 0x123456789ABCDEF0123456789ABCDEF] + o.rand_scalars(3, 77)
    bases = np.array([pr.to_words(g)] * len(exps), dtype=np.uint64)
    got = pairing.gt_pow(bases, h.scalars_to_np(exps))
    assert [pr.from_words(r) for r in got] == [pr.f12_pow(g, e) for e in exps]
    got_m = pairing.gt_pow(bases, h.scalars_to_np(exps, mont=True), mont=True)
    assert np.array_equal(got, got_m)


def test_multi_pairing_device_pointers(engine):
    """tb200_multi_pairing_dev on device-resident inputs equals the host-facing call."""
    n = 19
    ps, _ = o.rand_points(n, 41)
    qs, _ = o2.rand_points(n, 42)
    A, B = h.pts_to_np(ps), g2_np(qs)
    d = [ctypes.c_void_p() for _ in range(3)]
    for ptr, size in zip(d, (A.nbytes, B.nbytes, 576)):
        _lib.check(engine.tb200_dev_alloc(size, ctypes.byref(ptr)))
    try:
        _lib.check(engine.tb200_dev_upload(d[0], P(A), A.nbytes))
        _lib.check(engine.tb200_dev_upload(d[1], P(B), B.nbytes))
        _lib.check(engine.tb200_multi_pairing_dev(d[0], d[1], n, d[2], None))
        _lib.check(engine.tb200_stream_sync())
        out = np.zeros(72, dtype=np.uint64)
        _lib.check(engine.tb200_dev_download(P(out), d[2], 576))
    finally:
        for ptr in d:
            engine.tb200_dev_free(ptr)
    assert np.array_equal(out, pairing.multi_pairing(A, B))
    assert pr.from_words(out) == pr.multi_pairing(ps, qs)


def test_sharded_pairing_product_entry_points(engine):
    """multi_pairing(a, b) == final_exponentiation_of_product([miller_product(slice) ...]) for uneven slices (one of
    them empty, one beyond the warp-per-pair limit), and through parallel.multi_pairing_sharded at world size 1."""
    from testudo_b200 import parallel

    n = 2100
    ps, dl1 = o.rand_points(n, 51)
    qs, dl2 = o2.rand_points(n, 52)
    A, B = h.pts_to_np(ps), g2_np(qs)
    whole = pairing.multi_pairing(A, B)
    cuts = [0, 0, 3, 40, n]
    parts = [pairing.miller_product(A[lo:hi], B[lo:hi]) for lo, hi in zip(cuts[:-1], cuts[1:])]
    assert pr.from_words(parts[0]) == pr.F12_ONE
    assert np.array_equal(pairing.final_exponentiation_of_product(np.array(parts)), whole)
    assert np.array_equal(parallel.multi_pairing_sharded(A, B), whole)
    assert pr.from_words(whole) == pr.f12_pow(E_GEN, sum(a * b for a, b in zip(dl1, dl2)) % o.R_ORDER)
    # a partial value followed by the oracle's final exponentiation is the oracle's pairing product of that slice
    assert pr.final_exponentiation(pr.from_words(parts[2])) == pr.multi_pairing(ps[3:40], qs[3:40])


@pytest.mark.parametrize("n", [3, 600])
def test_pairing_team_sizes_agree(engine, n):
    """The cooperative kernels with a two-warp team, a one-warp team and the pipelined three-warp Miller kernel
    (tb200_set_pairing_team) give the same GT value; 600 pairs take the one-warp Miller kernel by default, 3 the pipelined."""
    gp = pairing
    ps, _ = o.rand_points(8, 71)
    qs, _ = o2.rand_points(8, 72)
    a = np.tile(np.array([o.affine_to_words(p) for p in ps], dtype=np.uint64).reshape(-1, 12), ((n + 7) // 8, 1))[:n].copy()
    b = np.tile(np.array([o2.affine_to_words(q) for q in qs], dtype=np.uint64).reshape(-1, 24), ((n + 7) // 8, 1))[:n].copy()
    outs = []
    for team in (64, 32, 33, 96, 0):
        engine.tb200_set_pairing_team(team)
        outs.append(gp.multi_pairing(a, b))
    engine.tb200_set_pairing_team(0)
    assert all(np.array_equal(outs[0], x) for x in outs[1:])
    if n == 3:
        assert pr.from_words(outs[0]) == pr.multi_pairing(ps[:3], qs[:3])
