"""GPU parity: single variable-base MSM through the C ABI vs golden vectors, the oracles and closed forms."""
import ctypes

import numpy as np
import pytest

import helpers as h
from oracle import bls12_377 as o
from testudo_b200 import _lib, msm

pytestmark = pytest.mark.gpu
GOLD = h.load_golden("msm_golden.json")


def P(a):
    return a.ctypes.data_as(ctypes.c_void_p)


@pytest.mark.parametrize("case", GOLD["seeded"], ids=lambda c: f"n{c['n']}")
def test_msm_seeded_golden(engine, case):
    pts, _ = o.rand_points(case["n"], case["points_seed"])
    sc = o.rand_scalars(case["n"], case["scalars_seed"])
    exp = h.pt_unhex(case["result"])
    assert h.pt_from_np(msm.msm_bigint(h.pts_to_np(pts), h.scalars_to_np(sc))) == exp
    assert h.pt_from_np(msm.msm_unchecked(h.pts_to_np(pts), h.scalars_to_np(sc, mont=True))) == exp


@pytest.mark.parametrize("mode", [0, 3])
@pytest.mark.parametrize("case", GOLD["explicit"] + GOLD["edge"], ids=lambda c: c.get("name", "explicit"))
def test_msm_explicit_and_edge_golden(engine, case, mode):
    pts = [h.pt_unhex(p) for p in case["points"]]
    sc = [int(s, 16) for s in case["scalars"]]
    engine.tb200_set_accumulate_mode(mode)
    try:
        got = msm.msm_bigint(h.pts_to_np(pts), h.scalars_to_np(sc))
    finally:
        engine.tb200_set_accumulate_mode(0)
    assert h.pt_from_np(got) == h.pt_unhex(case["result"])


def test_msm_empty_and_length_rules(engine):
    pts, _ = o.rand_points(4, 1)
    B = h.pts_to_np(pts)
    S = h.scalars_to_np([1, 2, 3], mont=True)
    assert msm.msm(B, S) == ("err", 3)                                    # VariableBaseMSM::msm -> Err(min_len)
    assert h.pt_from_np(msm.msm_unchecked(B, S)) == o.msm_naive(pts[:3], [1, 2, 3])  # truncates silently
    ok, val = msm.msm(B[:3], S)
    assert ok == "ok" and h.pt_from_np(val) == o.msm_naive(pts[:3], [1, 2, 3])
    assert h.pt_from_np(msm.msm_bigint(B[:0], S[:0])) is None            # n == 0 -> identity


@pytest.mark.parametrize("mode", [0, 3])
@pytest.mark.parametrize("c", [3, 5, 8, 11, 13, 16])
def test_msm_every_window_width(engine, oracle_c, c, mode):
    engine.tb200_set_accumulate_mode(mode)
    try:
        _every_window_width(engine, c)
    finally:
        engine.tb200_set_accumulate_mode(0)


def _every_window_width(engine, c):
    n = 700
    pts, dl = o.rand_points(n, 300 + c)
    sc = o.rand_scalars(n, 400 + c)
    sc[0] = o.R_ORDER - 1; sc[1] = 0; sc[2] = 1 << 252
    engine.tb200_set_window_bits(c)
    try:
        got = msm.msm_bigint(h.pts_to_np(pts), h.scalars_to_np(sc))
        cc = ctypes.c_int()
        engine.tb200_last_geometry(ctypes.byref(cc), None, None, None, None)
        assert cc.value == c
    finally:
        engine.tb200_set_window_bits(0)
    assert h.pt_from_np(got) == o.msm_by_dlog(dl, sc)


@pytest.mark.parametrize("logn", [10, 14, 16])
def test_msm_vs_c_oracle_and_dlog(engine, oracle_c, logn):
    n = 1 << logn
    a, step = 1234567 + logn, 0xABCDEF01
    bases = oracle_c.gen_points(h.pts_to_np([o.mul(a, o.G)])[0], h.pts_to_np([o.mul(step, o.G)])[0], n)
    sc = h.np_rand_scalars(n, logn)
    got = msm.msm_bigint(bases, sc)
    assert np.array_equal(got, oracle_c.msm_g1(bases, sc))
    ints = h.np_scalars_to_ints(sc)
    assert h.pt_from_np(got) == o.mul(sum(s * (a + step * k) for k, s in enumerate(ints)) % o.R_ORDER, o.G)


@pytest.mark.parametrize("mode", [0, 3])
def test_msm_skewed_scalars_heavy_buckets(engine, oracle_c, mode):
    engine.tb200_set_accumulate_mode(mode)
    try:
        _skewed_heavy_buckets(oracle_c)
    finally:
        engine.tb200_set_accumulate_mode(0)


def _skewed_heavy_buckets(oracle_c):
    """R1CS-like witness: 50% zeros, 25% ones, 25% uniform, repeated bases -> a few huge
    buckets that span many accumulation segments and exercise the head/fix-up path and the doubling branch."""
    n = 1 << 15
    bases = oracle_c.gen_points(h.pts_to_np([o.mul(99, o.G)])[0], h.pts_to_np([o.mul(31337, o.G)])[0], n)
    bases[1000:1200] = bases[1000]          # duplicates: P + P inside a bucket
    bases[5000:5010] = 0                    # identity bases
    sc = h.np_rand_scalars(n, 5)
    rng = np.random.default_rng(6)
    kind = rng.integers(0, 4, size=n)
    sc[kind < 2] = 0
    sc[kind == 2] = np.array([1, 0, 0, 0], dtype=np.uint64)
    got = msm.msm_bigint(bases, sc)
    assert np.array_equal(got, oracle_c.msm_g1(bases, sc))


def test_msm_linearity_property(engine, oracle_c):
    """MSM(B, s1) + MSM(B, s2) == MSM(B, s1 + s2 mod r) at 2^17 points (size-independent property)."""
    n = 1 << 17
    bases = oracle_c.gen_points(h.pts_to_np([o.mul(7, o.G)])[0], h.pts_to_np([o.mul(11, o.G)])[0], n)
    s1 = h.np_rand_scalars(n, 1)
    s2 = h.np_rand_scalars(n, 2)
    i1, i2 = h.np_scalars_to_ints(s1), h.np_scalars_to_ints(s2)
    s3 = h.scalars_to_np([(x + y) % o.R_ORDER for x, y in zip(i1, i2)])
    r1, r2, r3 = msm.msm_bigint(bases, s1), msm.msm_bigint(bases, s2), msm.msm_bigint(bases, s3)
    assert np.array_equal(msm.g1_sum(np.stack([r1, r2])), r3)
    assert h.pt_from_np(r3) == o.mul(sum((x + y) * (7 + 11 * k) for k, (x, y) in enumerate(zip(i1, i2))) % o.R_ORDER, o.G)


# ---- heavy / degenerate bucket contents through both accumulate variants (0 / 4: fused Y3, 3: plain CIOS products) ----------
@pytest.mark.parametrize("mode", [0, 3])
@pytest.mark.parametrize("c", [3, 6, 11, 16])
def test_skewed_duplicates_identities(engine, oracle_c, c, mode):
    """Heavy buckets, repeated bases (P + P inside a segment), P + (-P), identity bases, zero scalars."""
    n = 1 << 14
    bases = oracle_c.gen_points(h.pts_to_np([o.mul(99, o.G)])[0], h.pts_to_np([o.mul(31337, o.G)])[0], n)
    bases[1000:1200] = bases[1000]
    bases[3000:3100:2] = h.pts_to_np([o.neg(h.pt_from_np(bases[1000]))])[0]   # negations of a repeated base
    bases[5000:5010] = 0
    sc = h.np_rand_scalars(n, 50 + c)
    kind = np.random.default_rng(60 + c).integers(0, 4, size=n)
    sc[kind < 2] = 0
    sc[kind == 2] = np.array([1, 0, 0, 0], dtype=np.uint64)
    sc[1000:1200] = np.array([7, 0, 0, 0], dtype=np.uint64)
    sc[3000:3100:2] = np.array([7, 0, 0, 0], dtype=np.uint64)
    engine.tb200_set_window_bits(c)
    engine.tb200_set_accumulate_mode(mode)
    try:
        got = msm.msm_bigint(bases, sc)
    finally:
        engine.tb200_set_window_bits(0)
        engine.tb200_set_accumulate_mode(0)
    assert np.array_equal(got, oracle_c.msm_g1(bases, sc))


def test_accumulate_variants_agree_large(engine, oracle_c):
    n = 1 << 18
    bases = oracle_c.gen_points(h.pts_to_np([o.mul(3, o.G)])[0], h.pts_to_np([o.mul(5, o.G)])[0], n)
    sc = h.np_rand_scalars(n, 77)
    a = msm.msm_bigint(bases, sc)
    engine.tb200_set_accumulate_mode(3)
    try:
        c3 = msm.msm_bigint(bases, sc)
    finally:
        engine.tb200_set_accumulate_mode(0)
    assert np.array_equal(a, c3)
    ints = h.np_scalars_to_ints(sc)
    assert h.pt_from_np(a) == o.mul(sum(s * (3 + 5 * k) for k, s in enumerate(ints)) % o.R_ORDER, o.G)


# ---- small-n fast path (kernels_small.cuh: one CTA, a quad of lanes per point) -----------------------------------------------
@pytest.mark.parametrize("n", [1, 2, 3, 7, 8, 9, 31, 33, 63, 64, 65, 257, 1000, 1024])
def test_small_path_vs_oracle_and_pipeline(engine, oracle_c, n):
    """n <= 1024 takes the Straus path (8 points per one-warp CTA); tb200_set_small_msm_max(0) sends the same inputs through the sort pipeline: same
    points, both equal to the C oracle. Montgomery and canonical scalars."""
    bases = oracle_c.gen_points(h.pts_to_np([o.mul(1000 + n, o.G)])[0], h.pts_to_np([o.mul(77, o.G)])[0], n)
    sc = h.np_rand_scalars(n, 900 + n)
    want = oracle_c.msm_g1(bases, sc)
    assert np.array_equal(msm.msm_bigint(bases, sc), want)
    ints = h.np_scalars_to_ints(sc)
    assert np.array_equal(msm.msm_unchecked(bases, h.scalars_to_np(ints, mont=True)), want)
    engine.tb200_set_small_msm_max(0)
    try:
        assert np.array_equal(msm.msm_bigint(bases, sc), want)
    finally:
        engine.tb200_set_small_msm_max(-1)


def test_small_path_exceptional_cases(engine, oracle_c):
    """Identity bases, zero scalars, r - 1, 2^252, the same point many times (partial sums that coincide: the tree's
    doubling branch), P and -P with equal scalars (sum = identity), scalars whose radix-16 digits are all +-8 / 0."""
    r = o.R_ORDER
    P1, P2 = o.mul(5, o.G), o.mul(11, o.G)
    cases = [
        ([P1, P1], [7, 7]),
        ([P1, o.neg(P1)], [9, 9]),
        ([P1, o.neg(P1), P2], [9, 9, 3]),
        ([None, P1, None], [5, 6, 7]),
        ([P1, P2], [0, 0]),
        ([P1, P2, P1], [r - 1, 1 << 252, r - 2]),
        ([P1] * 64, [3] * 64),
        ([P1] * 5, [0x8888888888888888, 0x0808080808080808, 8, 0x80, (1 << 252) + 0x88]),
        ([o.G], [1]),
        ([P2], [r - 1]),
    ]
    for pts, sc in cases:
        got = msm.msm_bigint(h.pts_to_np(pts), h.scalars_to_np(sc))
        assert h.pt_from_np(got) == o.msm_naive(pts, sc), (pts, sc)


def test_msm_in_flight_next_to_other_calls(engine, oracle_c):
    """tb200_msm_g1_begin / _end: MSMs of 0, 5, 1000 (Straus path) and 2^14 points (sort pipeline on a side workspace) started,
    OTHER library calls made while they are in flight (an MSM through the main pipeline, a second job), then collected:
    every result equals the C oracle's and the blocking entry point's; a discarded job (_end(job, NULL)) leaves the library
    usable."""
    import ctypes

    jobs = []
    for n in (0, 5, 1000, 1 << 14):
        bases = oracle_c.gen_points(h.pts_to_np([o.mul(4000 + n, o.G)])[0], h.pts_to_np([o.mul(91, o.G)])[0], max(n, 1))[:n]
        sc = h.scalars_to_np(h.np_scalars_to_ints(h.np_rand_scalars(max(n, 1), 1700 + n))[:n], mont=True).reshape(n, 4)
        jobs.append((n, bases, sc, msm.msm_unchecked_begin(bases, sc)))
    n2 = 1 << 15
    other_b = oracle_c.gen_points(h.pts_to_np([o.mul(5, o.G)])[0], h.pts_to_np([o.mul(7, o.G)])[0], n2)
    other_s = h.np_rand_scalars(n2, 1800)
    assert np.array_equal(msm.msm_bigint(other_b, other_s), oracle_c.msm_g1(other_b, other_s))   # while the jobs run
    for n, bases, sc, job in jobs:
        got = job.wait()
        want = msm.msm_unchecked(bases, sc) if n else np.zeros(12, dtype=np.uint64)
        assert np.array_equal(got, want), n
        if n:
            canon = h.scalars_to_np(h.scalars_from_np(sc, mont=True))
            assert np.array_equal(got, oracle_c.msm_g1(bases, canon)), n
    handle = ctypes.c_void_p()
    P = lambda a: a.ctypes.data_as(ctypes.c_void_p)   # noqa: E731
    assert engine.tb200_msm_g1_begin(P(other_b), P(other_s), n2, 0, ctypes.byref(handle)) == 0
    assert engine.tb200_msm_g1_end(handle, None) == 0                                              # discarded
    assert engine.tb200_msm_g1_begin(None, None, 3, 0, ctypes.byref(handle)) == -1
    assert engine.tb200_msm_g1_end(None, None) == -1
    assert np.array_equal(msm.msm_bigint(other_b, other_s), oracle_c.msm_g1(other_b, other_s))


def test_ragged_row_batch_vs_c_oracle(engine, oracle_c):
    """tb200_msm_g1_rows: independent MSMs of 0, 1, 2, 8, 9, 27, 13, 64, 1000 and 1024 points in ONE launch (rows above 8
    points span several CTAs and are summed by the last one to finish, one ticket per row) -- every row equals the C oracle's
    MSM of its slice; an empty row is the identity; a row of 1025 points is refused."""
    import ctypes

    lens = [0, 1, 2, 8, 9, 27, 13, 64, 1000, 0, 1024, 3]
    n = sum(lens)
    bases = oracle_c.gen_points(h.pts_to_np([o.mul(31337, o.G)])[0], h.pts_to_np([o.mul(1009, o.G)])[0], n)
    sc = h.np_rand_scalars(n, 2222)
    sc[5] = 0                                               # a zero scalar
    bases[7] = 0                                            # an identity base
    got = msm.msm_rows(bases, sc, lens)
    assert got.shape == (len(lens), 12)
    pos = 0
    for i, ln in enumerate(lens):
        want = oracle_c.msm_g1(bases[pos:pos + ln], sc[pos:pos + ln]) if ln else np.zeros(12, dtype=np.uint64)
        assert np.array_equal(got[i], want), (i, ln)
        pos += ln
    batch = msm.RowBatch()
    r0 = batch.add(bases[:5], sc[:5])
    r1 = batch.add(bases[5:40], sc[5:40])
    pts = batch.run()
    assert np.array_equal(pts[r0], oracle_c.msm_g1(bases[:5], sc[:5])) and np.array_equal(pts[r1], oracle_c.msm_g1(bases[5:40], sc[5:40]))
    with pytest.raises(ValueError):
        msm.msm_rows(bases[:4], sc[:4], [1, 2])
    too_long = np.array([1025], dtype=np.uint64)
    P = lambda a: a.ctypes.data_as(ctypes.c_void_p)   # noqa: E731
    assert engine.tb200_msm_g1_rows(P(bases), P(sc), P(too_long), 1, 0, P(np.zeros(12, dtype=np.uint64))) == -3
