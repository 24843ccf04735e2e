"""GPU parity: MIPP G1 steps (cross commitments, compress, device-resident prover loop) vs the oracles.
Follows src/mipp.rs:58-120 line by line on the CPU with the big-int oracle and compares every value."""
import hashlib

import numpy as np
import pytest

import helpers as h
from oracle import bls12_377 as o
from testudo_b200 import _lib, mipp

pytestmark = pytest.mark.gpu


def make_challenge():
    state = hashlib.sha256(b"mipp-test")

    def challenge(label, points):
        state.update(label)
        for p in points:
            state.update(np.asarray(p, dtype=np.uint64).tobytes())
        return int.from_bytes(state.digest(), "little") % o.R_ORDER or 1

    return challenge


def test_multiexponentiation_error_behaviour(engine):
    pts, _ = o.rand_points(4, 1)
    with pytest.raises(mipp.InvalidIPVectorLength):          # src/mipp.rs:389-391
        mipp.multiexponentiation(h.pts_to_np(pts), h.scalars_to_np([1, 2, 3], mont=True))
    got = mipp.multiexponentiation(h.pts_to_np(pts), h.scalars_to_np([1, 2, 3, 4], mont=True))
    assert h.pt_from_np(got) == o.msm_naive(pts, [1, 2, 3, 4])


@pytest.mark.parametrize("n", [2, 8, 64])
def test_compress_matches_oracle(engine, oracle_c, n):
    pts, _ = o.rand_points(n, 50 + n)
    if n >= 8:
        pts[1] = None                   # identity on the left
        pts[n // 2 + 2] = None          # identity on the right
        pts[n // 2 + 3] = pts[3]        # c * P + P
    k = o.rand_scalars(1, 60 + n)[0]
    got = mipp.compress(h.pts_to_np(pts), n // 2, h.scalars_to_np([k], mont=True)[0])
    exp = oracle_c.compress_g1(h.pts_to_np(pts), n // 2, h.scalars_to_np([k])[0])
    assert np.array_equal(got, exp)
    for i in range(n // 2):
        assert h.pt_from_np(got[i]) == o.add(pts[i], o.mul(k, pts[n // 2 + i]))


@pytest.mark.parametrize("n", [2, 16, 128])
def test_mipp_prover_loop_matches_reference_semantics(engine, n):
    pts, _ = o.rand_points(n, 70 + n)
    if n >= 16:
        pts[5] = None  # identity commitments appear for all-zero rows
    y = o.rand_scalars(n, 80 + n)
    U = h.pts_to_np([o.msm_naive(pts, y)])[0]
    proof = mipp.MippProofG1.prove(make_challenge(), h.pts_to_np(pts), h.scalars_to_np(y, mont=True), U)
    # CPU restatement of the loop with the same challenges
    ch = make_challenge()
    ch(b"U", [U])
    m_a, m_y = list(pts), list(y)
    rounds = 0
    while len(m_a) > 1:
        split = len(m_a) // 2
        a_l, a_r, y_l, y_r = m_a[:split], m_a[split:], m_y[:split], m_y[split:]
        u_l = o.msm_naive(a_l, y_r)                      # multiexponentiation(ra_l, &ry_r)  src/mipp.rs:82
        u_r = o.msm_naive(a_r, y_l)                      # multiexponentiation(ra_r, &ry_l)  src/mipp.rs:84
        assert h.pt_from_np(proof.comms_u[rounds][0]) == u_l
        assert h.pt_from_np(proof.comms_u[rounds][1]) == u_r
        c_inv = ch(b"challenge_i", [h.pts_to_np([u_l])[0], h.pts_to_np([u_r])[0]])
        c = pow(c_inv, -1, o.R_ORDER)
        assert proof.xs[rounds] == c and proof.xs_inv[rounds] == c_inv
        m_a = [o.add(a_l[i], o.mul(c, a_r[i])) for i in range(split)]          # compress, src/mipp.rs:354-367
        m_y = [(y_l[i] + c_inv * y_r[i]) % o.R_ORDER for i in range(split)]    # compress_field, :370-383
        rounds += 1
    assert rounds == n.bit_length() - 1
    assert h.pt_from_np(proof.final_a) == m_a[0]
    assert h.scalars_from_np(proof.final_y, mont=True)[0] == m_y[0]


def test_mipp_round_entry_points_agree(engine):
    """One MIPP round through the fused call (tb200_mipp_cross_all) and through the separate calls
    (tb200_mipp_g1_cross + tb200_mipp_pairing_cross), before and after a fold, against the oracle's cross
    commitments and pairing products (src/mipp.rs:77-94)."""
    import ctypes

    from oracle import bls12_377_g2 as o2
    from oracle import pairing as pr

    def P(a):
        return a.ctypes.data_as(ctypes.c_void_p)

    n = 8
    a, _ = o.rand_points(n, 61)
    hk, _ = o2.rand_points(n, 62)
    y = o.rand_scalars(n, 63)
    A = h.pts_to_np(a)
    H = np.array([o2.affine_to_words(q) for q in hk], dtype=np.uint64)
    Y = h.scalars_to_np(y, mont=True)
    ha, hh = ctypes.c_void_p(), ctypes.c_void_p()
    _lib.check(engine.tb200_mipp_g1_begin(P(A), P(Y), n, _lib.SCALARS_MONT, ctypes.byref(ha)))
    _lib.check(engine.tb200_mipp_g2_begin(P(H), n, _lib.SCALARS_MONT, ctypes.byref(hh)))
    try:
        m_a, m_h, m_y = list(a), list(hk), list(y)
        for rnd in range(2):
            split = len(m_a) // 2
            ul, ur = np.zeros(12, np.uint64), np.zeros(12, np.uint64)
            tl, tr = np.zeros(72, np.uint64), np.zeros(72, np.uint64)
            _lib.check(engine.tb200_mipp_cross_all(ha, hh, P(ul), P(ur), P(tl), P(tr)))
            ul2, ur2 = np.zeros(12, np.uint64), np.zeros(12, np.uint64)
            tl2, tr2 = np.zeros(72, np.uint64), np.zeros(72, np.uint64)
            _lib.check(engine.tb200_mipp_g1_cross(ha, P(ul2), P(ur2)))
            _lib.check(engine.tb200_mipp_pairing_cross(ha, hh, P(tl2), P(tr2)))
            assert np.array_equal(ul, ul2) and np.array_equal(ur, ur2)
            assert np.array_equal(tl, tl2) and np.array_equal(tr, tr2)
            assert h.pt_from_np(ul) == o.msm_naive(m_a[:split], m_y[split:])
            assert h.pt_from_np(ur) == o.msm_naive(m_a[split:], m_y[:split])
            assert pr.from_words(tl) == pr.multi_pairing(m_a[:split], m_h[split:])
            assert pr.from_words(tr) == pr.multi_pairing(m_a[split:], m_h[:split])
            c_inv = o.rand_scalars(1, 64 + rnd)[0]
            c = pow(c_inv, -1, o.R_ORDER)
            _lib.check(engine.tb200_mipp_g1_fold(ha, P(h.scalars_to_np([c], mont=True)), P(h.scalars_to_np([c_inv], mont=True))))
            _lib.check(engine.tb200_mipp_g2_fold(hh, P(h.scalars_to_np([c_inv], mont=True))))
            m_a = [o.add(m_a[i], o.mul(c, m_a[split + i])) for i in range(split)]
            m_y = [(m_y[i] + c_inv * m_y[split + i]) % o.R_ORDER for i in range(split)]
            m_h = [o2.add(m_h[i], o2.mul(c_inv, m_h[split + i])) for i in range(split)]
        fh = np.zeros((2, 24), np.uint64)
        _lib.check(engine.tb200_mipp_g2_read(hh, P(fh)))
        assert [o2.affine_from_words(r) for r in fh] == m_h          # the twisted-Frobenius fold, bit-exact
    finally:
        engine.tb200_mipp_g1_end(ha)
        engine.tb200_mipp_g2_end(hh)


@pytest.mark.parametrize("m", [0, 1, 2, 5, 13])
def test_structured_polynomial_on_device_equals_the_reference_loop(engine, m):
    """`polynomial_evaluations_from_transcript` (src/mipp.rs:159-180) computed by tb200_fr_subset_products, in ark's
    Montgomery form, against the integer loop."""
    from testudo_b200 import fr

    cs_inv = o.rand_scalars(m, 3100 + m)
    want = mipp.polynomial_evaluations_from_transcript(cs_inv)
    got = fr.from_mont_words(mipp.polynomial_evaluations_words(cs_inv))
    assert got == want
    if m:
        assert want[1] == cs_inv[m - 1] and want[1 << (m - 1)] == cs_inv[0]      # bit j from the lsb selects cs_inv[m-j-1]
