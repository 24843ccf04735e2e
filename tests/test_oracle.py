"""CPU tests: the oracle (Python big-int + C restatement of ark's Pippenger) against the golden vectors."""
import numpy as np
import pytest

import helpers as h
from oracle import bls12_377 as o

GOLD = h.load_golden("msm_golden.json")


def test_curve_constants_and_kats():
    assert o.is_on_curve(o.G)
    assert o.mul(o.R_ORDER, o.G) is None
    k = GOLD["curve_kats"]
    assert h.pt_unhex(k["G"]) == o.G
    assert h.pt_unhex(k["2G"]) == o.add(o.G, o.G)
    assert h.pt_unhex(k["3G"]) == o.mul(3, o.G)
    assert h.pt_unhex(k["(r-1)G"]) == o.neg(o.G)
    assert h.pt_unhex(k["12345G"]) == o.mul(12345, o.G)
    # Montgomery constants of SURVEY.md App. B
    assert o.FQ_R == 0x8D6661E2FDF49A4CF495BF803C84E87B4E97B76E7C63059F7DB3A98A7D3FF251409F837FFFFFB102CDFFFFFFFFFF68
    assert (-pow(o.Q, -1, 1 << 64)) % (1 << 64) == 0x8508BFFFFFFFFFFF
    assert (-pow(o.R_ORDER, -1, 1 << 64)) % (1 << 64) == 0x0A117FFFFFFFFFFF


def test_ark_window_rule(oracle_c):
    # SURVEY.md App. A.1 table
    for n, c in ((1 << 10, 8), (1 << 12, 10), (1 << 13, 10), (1 << 20, 15), (1 << 24, 18), (1 << 26, 19), (31, 3)):
        assert o.ark_window_bits(n) == c
        assert oracle_c.ark_window_bits(n) == c


def test_ark_digits_recompose():
    for s in [0, 1, o.R_ORDER - 1, 1 << 252] + o.rand_scalars(50, 7):
        for w in (3, 8, 10, 15, 18):
            d = o.ark_make_digits(s, w)
            assert sum(v << (w * i) for i, v in enumerate(d)) == s


@pytest.mark.parametrize("case", GOLD["seeded"], ids=lambda c: f"n{c['n']}")
def test_c_oracle_seeded_golden(oracle_c, case):
    pts, _ = o.rand_points(case["n"], case["points_seed"])
    sc = o.rand_scalars(case["n"], case["scalars_seed"])
    got = h.pt_from_np(oracle_c.msm_g1(h.pts_to_np(pts), h.scalars_to_np(sc)))
    assert got == h.pt_unhex(case["result"])
    got_m = h.pt_from_np(oracle_c.msm_g1(h.pts_to_np(pts), h.scalars_to_np(sc, mont=True), mont=True))
    assert got_m == h.pt_unhex(case["result"])


@pytest.mark.parametrize("case", GOLD["explicit"] + GOLD["edge"], ids=lambda c: c.get("name", "explicit"))
def test_oracles_explicit_and_edge_golden(oracle_c, case):
    pts = [h.pt_unhex(p) for p in case["points"]]
    sc = [int(s, 16) for s in case["scalars"]]
    exp = h.pt_unhex(case["result"])
    assert o.msm_naive(pts, sc) == exp
    assert o.msm_pippenger(pts, sc) == exp
    assert h.pt_from_np(oracle_c.msm_g1(h.pts_to_np(pts), h.scalars_to_np(sc))) == exp


def test_msm_length_rules():
    pts, _ = o.rand_points(4, 1)
    assert o.msm_checked(pts, [1, 2, 3]) == ("err", 3)  # VariableBaseMSM::msm -> Err(min_len)
    assert o.msm_checked(pts, [1, 2, 3, 4])[0] == "ok"
    assert o.msm_naive(pts, [1, 2]) == o.msm_naive(pts[:2], [1, 2])  # msm_unchecked truncates


@pytest.mark.parametrize("case", GOLD["sqrt_rows"], ids=lambda c: f"nv{c['num_vars']}")
def test_c_oracle_sqrt_rows(oracle_c, case):
    nv = case["num_vars"]
    m_col = nv // 2
    m_row = nv - m_col
    srs, _ = o.rand_points(1 << m_row, case["srs_seed"])
    z = o.rand_scalars(1 << nv, case["z_seed"])
    out = oracle_c.msm_g1_batch(h.pts_to_np(srs), h.scalars_to_np(z), 1 << m_col, 1 << m_row, 1, 1 << m_col)
    assert [h.pt_from_np(r) for r in out] == [h.pt_unhex(r) for r in case["rows"]]


def test_c_oracle_compress_and_group(oracle_c):
    pts, _ = o.rand_points(8, 77)
    k = o.rand_scalars(1, 78)[0]
    got = oracle_c.compress_g1(h.pts_to_np(pts), 4, h.scalars_to_np([k])[0])
    for i in range(4):  # src/mipp.rs:354-367
        assert h.pt_from_np(got[i]) == o.add(pts[i], o.mul(k, pts[4 + i]))
    assert h.pt_from_np(oracle_c.g1_add(h.pts_to_np([pts[0]])[0], h.pts_to_np([pts[0]])[0])) == o.add(pts[0], pts[0])
    assert h.pt_from_np(oracle_c.g1_mul(h.pts_to_np([pts[1]])[0], h.scalars_to_np([k])[0])) == o.mul(k, pts[1])


def test_c_oracle_large_by_dlog(oracle_c):
    """2^14 points through the C Pippenger vs the closed-form discrete-log oracle."""
    n = 1 << 14
    a, b, step = 123456789, 987654321, 0xABCDEF
    start = h.pts_to_np([o.mul(a, o.G)])[0]
    stp = h.pts_to_np([o.mul(step, o.G)])[0]
    bases = oracle_c.gen_points(start, stp, n)
    sc = h.np_rand_scalars(n, 99)
    ints = h.np_scalars_to_ints(sc)
    exp = o.mul(sum(s * (a + step * k) for k, s in enumerate(ints)) % o.R_ORDER, o.G)
    assert h.pt_from_np(oracle_c.msm_g1(bases, sc)) == exp


def test_pst_open_restatement_satisfies_the_pst_identity():
    """oracle/pst.py (ark-poly-commit `MultilinearPC::open` restated): with a known trapdoor t the proofs' exponents
    satisfy f(t) - f(point) = sum_i (t_i - point_i) q_i(t_{i+1..}) -- the relation `check` verifies with pairings --
    and the MSM over the synthetic CRS level equals q_i(t_{i+1..}) * G."""
    from oracle import pst

    nv = 5
    t = o.rand_scalars(nv, 1)
    evals = o.rand_scalars(1 << nv, 2)
    point = o.rand_scalars(nv, 3)
    qs = pst.quotients(evals, point)
    dl = [pst.mle_eval(q, t[i + 1:]) for i, q in enumerate(qs)]
    lhs = (pst.mle_eval(evals, t) - pst.mle_eval(evals, point)) % o.R_ORDER
    assert lhs == sum((t[i] - point[i]) * dl[i] for i in range(nv)) % o.R_ORDER
    levels = [[o.mul(e, o.G) for e in pst.eq_exponents(t[k:])] for k in range(nv)]
    proofs = pst.open_proofs(evals, point, levels, o.msm_naive)
    assert proofs == [o.mul(d, o.G) for d in dl]
