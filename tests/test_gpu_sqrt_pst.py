"""GPU parity: sqrt_pst commit (shared-SRS batched MSM) and the G1 work of open, vs golden rows and the oracles.
Mirrors the reference's own tests check_sqrt_poly_eval / check_sqrt_poly_commit (src/sqrt_pst.rs:277-342) as far as
the G1 scope goes (no pairings): eval identity, commit -> open with the debug_assert identity of src/sqrt_pst.rs:206."""
import ctypes
import hashlib

import numpy as np
import pytest

import helpers as h
from oracle import bls12_377 as o
from testudo_b200 import _lib, commitments, fr, mipp, sqrt_pst

pytestmark = pytest.mark.gpu
GOLD = h.load_golden("msm_golden.json")


def P(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def fake_transcript():
    """Deterministic stand-in for the Poseidon transcript (out of scope): hashes what the prover appends."""
    state = hashlib.sha256(b"testudo-b200-test")

    def challenge(label, points):
        state.update(label)
        for p in points:
            state.update(np.asarray(p, dtype=np.uint64).tobytes())
        return int.from_bytes(state.digest(), "little") % o.R_ORDER or 1

    return challenge


@pytest.mark.parametrize("case", GOLD["sqrt_rows"], ids=lambda c: f"nv{c['num_vars']}")
def test_commit_rows_golden(engine, case):
    nv = case["num_vars"]
    m_row = nv - nv // 2
    srs, _ = o.rand_points(1 << m_row, case["srs_seed"])
    z = o.rand_scalars(1 << nv, case["z_seed"])
    ck = sqrt_pst.CommitterKey.from_points(h.pts_to_np(srs))
    poly = sqrt_pst.Polynomial.from_evaluations(h.scalars_to_np(z, mont=True))
    comm_list, t = poly.commit(ck)
    assert t is None
    assert [h.pt_from_np(r) for r in comm_list] == [h.pt_unhex(r) for r in case["rows"]]
    ck.close()


@pytest.mark.parametrize("mode", [0, 3])
@pytest.mark.parametrize("nv,window", [(12, 0), (13, 0), (16, 0), (17, 0), (12, 5), (14, 11)])
def test_commit_vs_c_oracle(engine, oracle_c, nv, window, mode):
    engine.tb200_set_accumulate_mode(mode)
    try:
        _commit_vs_c_oracle(engine, oracle_c, nv, window)
    finally:
        engine.tb200_set_accumulate_mode(0)


def _commit_vs_c_oracle(engine, oracle_c, nv, window):
    m_col = nv // 2
    m_row = nv - m_col
    srs = oracle_c.gen_points(h.pts_to_np([o.mul(31 + nv, o.G)])[0], h.pts_to_np([o.mul(977, o.G)])[0], 1 << m_row)
    z = h.np_rand_scalars(1 << nv, nv)
    z[5] = 0; z[6] = np.array(o.to_limbs64(o.R_ORDER - 1, 4), dtype=np.uint64); z[7] = np.array([1, 0, 0, 0], dtype=np.uint64)
    if nv == 12:
        z[3::64] = 0          # an all-zero row -> identity commitment (zero-padded witness)
    ck = sqrt_pst.CommitterKey.from_points(srs, window_bits=window)
    lib = _lib.engine()
    out = np.zeros((1 << m_col, 12), dtype=np.uint64)
    _lib.check(lib.tb200_msm_g1_batch(ck._h, P(z), 1 << m_col, 1 << m_row, 1, 1 << m_col, 0, P(out)))
    exp = oracle_c.msm_g1_batch(srs, z, 1 << m_col, 1 << m_row, 1, 1 << m_col)
    assert np.array_equal(out, exp)
    if nv == 12:
        assert not out[3].any()
    # the same rows handed over as separate heap buffers (what Polynomial::commit holds, src/sqrt_pst.rs:48-62)
    rows = [np.ascontiguousarray(z[i :: 1 << m_col]) for i in range(1 << m_col)]
    ptrs = (ctypes.c_void_p * len(rows))(*[r.ctypes.data for r in rows])
    out2 = np.zeros_like(out)
    _lib.check(lib.tb200_msm_g1_batch_ptrs(ck._h, ptrs, len(rows), 1 << m_row, 0, P(out2)))
    assert np.array_equal(out2, exp)
    ck.close()


def test_commit_closed_form_g_pow_poly_at_t(engine):
    """SRS g^{eq(t,x)} (ark-poly-commit setup): every row commitment is g^{p_i(t)}."""
    nv = 10
    m_col = m_row = 5
    t = o.rand_scalars(m_row, 555)
    eq = []
    for x in range(1 << m_row):   # little-endian variable order
        v = 1
        for k in range(m_row):
            v = v * (t[k] if (x >> k) & 1 else (1 - t[k])) % o.R_ORDER
        eq.append(v)
    srs = [o.mul(e, o.G) for e in eq]
    z = o.rand_scalars(1 << nv, 556)
    ck = sqrt_pst.CommitterKey.from_points(h.pts_to_np(srs))
    poly = sqrt_pst.Polynomial.from_evaluations(h.scalars_to_np(z, mont=True))
    comm_list, _ = poly.commit(ck)
    for i in range(1 << m_col):
        pit = sum(z[(j << m_col) | i] * eq[j] for j in range(1 << m_row)) % o.R_ORDER
        assert h.pt_from_np(comm_list[i]) == o.mul(pit, o.G)
    ck.close()


@pytest.mark.parametrize("nv", [5, 6])
def test_sqrt_eval_and_open_identities(engine, nv):
    """check_sqrt_poly_eval (src/sqrt_pst.rs:277-295) + the G1 chain of check_sqrt_poly_commit (:297-342)."""
    m_col = nv // 2
    m_row = nv - m_col
    z = o.rand_scalars(1 << nv, 600 + nv)
    r = o.rand_scalars(nv, 700 + nv)
    srs, _ = o.rand_points(1 << m_row, 800 + nv)
    poly = sqrt_pst.Polynomial.from_evaluations(h.scalars_to_np(z, mont=True))
    # DensePolynomial::evaluate (MSB-first variable order, src/dense_mlpoly.rs) == sqrt-layout eval
    direct = sum(zi * fr.get_chi_i(r, i) for i, zi in enumerate(z)) % o.R_ORDER
    assert poly.eval(r) == direct
    ck = sqrt_pst.CommitterKey.from_points(h.pts_to_np(srs))
    comm_list, _ = poly.commit(ck)
    opened = poly.open(fake_transcript(), comm_list, ck, r)          # asserts c_u == commit(q) inside
    q = fr.from_mont_words(poly.q)
    assert h.pt_from_np(opened.u) == o.msm_naive(srs, q)
    chis = fr.from_mont_words(poly.chis_b)
    assert h.pt_from_np(opened.u) == o.msm_naive([h.pt_from_np(c) for c in comm_list], chis)
    assert len(opened.mipp.comms_u) == m_col
    ck.close()


def test_hyrax_commit_inner_contiguous_rows(engine, oracle_c):
    """DensePolynomial::commit_inner (src/dense_mlpoly.rs:315-329): contiguous rows over shared gens, blinds 0 and != 0."""
    L, R = 32, 64
    G = oracle_c.gen_points(h.pts_to_np([o.mul(5, o.G)])[0], h.pts_to_np([o.mul(13, o.G)])[0], R)
    hpt = h.pts_to_np([o.mul(424242, o.G)])[0]
    z = h.np_rand_scalars(L * R, 9)
    z[:R] = 0
    z[R: 2 * R, 1:] = 0  # small scalars (addresses / timestamps)
    zm = h.scalars_to_np(h.np_scalars_to_ints(z), mont=True)
    gens = commitments.MultiCommitGens(G, hpt)
    rows = commitments.commit_inner(zm, np.zeros((L, 4), dtype=np.uint64), gens)
    exp = oracle_c.msm_g1_batch(G, z, L, R, R, 1)
    assert np.array_equal(rows, exp)
    assert not rows[0].any()
    blinds = o.rand_scalars(L, 10)
    rows_b = commitments.commit_inner(zm, h.scalars_to_np(blinds, mont=True), gens)
    for i in (0, 1, 17):
        assert h.pt_from_np(rows_b[i]) == o.add(h.pt_from_np(exp[i]), o.mul(blinds[i], h.pt_from_np(hpt)))
    # commit_slice / commit_scalar (src/commitments.rs:70-86)
    cs = commitments.PedersenCommit.commit_slice(zm[2 * R: 3 * R], h.scalars_to_np([blinds[2]], mont=True)[0], gens)
    assert np.array_equal(cs, rows_b[2])
    g1 = commitments.MultiCommitGens(G[:1], hpt)
    sc = commitments.PedersenCommit.commit_scalar(h.scalars_to_np([77], mont=True)[0], h.scalars_to_np([88], mont=True)[0], g1)
    assert h.pt_from_np(sc) == o.add(o.mul(77, h.pt_from_np(G[0])), o.mul(88, h.pt_from_np(hpt)))
    gens.close()


@pytest.mark.parametrize("nv", [1, 2, 7, 10, 13])
def test_get_q_chis_eval_on_device(engine, nv):
    """k_fr_chis / k_fr_matvec vs the reference's CPU loops (src/sqrt_pst.rs:81-115) in integers."""
    z = o.rand_scalars(1 << nv, 1000 + nv)
    z[0] = 0
    z[-1] = o.R_ORDER - 1
    r = o.rand_scalars(nv, 1100 + nv)
    poly = sqrt_pst.Polynomial.from_evaluations(h.scalars_to_np(z, mont=True))
    poly.get_q(r)
    q_host, chis_host = poly.get_q_host(r)
    assert np.array_equal(poly.q, q_host)
    assert np.array_equal(poly.chis_b, chis_host)
    direct = sum(zi * fr.get_chi_i(r, i) for i, zi in enumerate(z)) % o.R_ORDER      # check_sqrt_poly_eval
    assert poly.eval(r) == direct


def _crs_levels(engine, t, g2):
    """Synthetic CRS with known trapdoor: powers[k][x] = eq((t_k..), x) * generator."""
    from oracle import bls12_377_g2 as o2
    from oracle import pst

    levels = []
    for k in range(len(t)):
        e = pst.eq_exponents(t[k:])
        out = np.zeros((len(e), 24 if g2 else 12), np.uint64)
        gen = np.array([o2.affine_to_words(o2.G2)], dtype=np.uint64) if g2 else h.pts_to_np([o.G])
        gen = np.ascontiguousarray(np.tile(gen, (len(e), 1)))
        fn = engine.tb200_test_g2_mul if g2 else engine.tb200_test_g1_mul
        _lib.check(fn(P(gen), P(h.scalars_to_np(e)), len(e), P(out)))
        levels.append(out)
    return levels


@pytest.mark.parametrize("nv", [4, 7])
def test_open_with_full_crs_g2_proof_and_mipp_open_g1(engine, nv):
    """`Polynomial::open` with the whole CommitterKey (src/sqrt_pst.rs:168-230): the PST proof of q at a_rev (G2 MSMs,
    :225) and MIPP's final_h / pst_proof_h (src/mipp.rs:114-144) against the closed forms a known trapdoor gives."""
    from oracle import bls12_377_g2 as o2
    from oracle import pst

    m_col = nv // 2
    m_row = nv - m_col
    odd = nv % 2
    t = o.rand_scalars(m_row, 900 + nv)
    g_levels = _crs_levels(engine, t, False)
    h_levels = _crs_levels(engine, t, True)
    z = o.rand_scalars(1 << nv, 910 + nv)
    r = o.rand_scalars(nv, 920 + nv)
    poly = sqrt_pst.Polynomial.from_evaluations(h.scalars_to_np(z, mont=True))
    ck = sqrt_pst.CommitterKey.from_points(g_levels[0]).with_levels(g_levels, h_levels)
    comm_list, _ = poly.commit(ck)
    opened = poly.open(fake_transcript(), comm_list, ck, r)
    # PST proof of q at a_rev: proof_i = q_i(t_{i+1..}) * G2
    q = fr.from_mont_words(poly.q)
    a_rev = list(r[: m_col + odd])[::-1]
    dl = [pst.mle_eval(qi, t[i + 1:]) for i, qi in enumerate(pst.quotients(q, a_rev))]
    assert [o2.affine_from_words(p) for p in opened.pst_proof] == [o2.mul(d, o2.G2) for d in dl]
    # and U = q(t) * G closes the PST identity against the claimed evaluation q(a_rev)
    assert h.pt_from_np(opened.u) == o.mul(pst.mle_eval(q, t), o.G)
    assert (pst.mle_eval(q, t) - pst.mle_eval(q, a_rev)) % o.R_ORDER == \
        sum((t[i] - a_rev[i]) * dl[i] for i in range(m_row)) % o.R_ORDER
    # MIPP: final_h = p_h(t_odd..) * G2 and the open_g1 proof of p_h at rs over powers_of_g[odd + i]
    mp = opened.mipp
    ph = mipp.polynomial_evaluations_from_transcript(mp.xs_inv)
    assert o2.affine_from_words(mp.final_h) == o2.mul(pst.mle_eval(ph, t[odd:]), o2.G2)
    dlh = [pst.mle_eval(qi, t[odd + i + 1:]) for i, qi in enumerate(pst.quotients(ph, mp.rs))]
    assert [h.pt_from_np(p) for p in mp.pst_proof_h] == [o.mul(d, o.G) for d in dlh]
    ck.close()


def _neutral(values):
    """Encodes what is appended to the transcript as the C-ABI word arrays, for prover (numpy) and verifier (oracle)."""
    from oracle import bls12_377_g2 as o2
    from oracle import pairing as pr

    out = []
    for v in values:
        if isinstance(v, tuple) and v and v[0] in ("g1", "g2", "gt"):
            kind, val = v
            words = o.affine_to_words(val) if kind == "g1" else (o2.affine_to_words(val) if kind == "g2" else pr.to_words(val))
            out.append(np.array(words, dtype=np.uint64))
        else:
            out.append(np.asarray(v, dtype=np.uint64))
    return out


def shared_transcript():
    inner = fake_transcript()
    return lambda label, values: inner(label, _neutral(values))


@pytest.mark.parametrize("nv", [5, 6])
def test_check_sqrt_poly_commit_roundtrip_with_oracle_verifier(engine, nv):
    """The reference's own test for this path, `check_sqrt_poly_commit(5)` / `(6)` (src/sqrt_pst.rs:297-342): setup,
    commit (comm_list AND the pairing product t), open (U, PST proof, MIPP proof with comms_t), verify == true -- with
    every prover value produced by the GPU path and the verifier being the independent big-integer restatement
    (oracle/verifier.py). Tampered proofs must be rejected."""
    from oracle import bls12_377_g2 as o2
    from oracle import pairing as pr
    from oracle import verifier as ver

    m_col = nv // 2
    m_row = nv - m_col
    odd = nv % 2
    t = o.rand_scalars(m_row, 1900 + nv)
    g_levels = _crs_levels(engine, t, False)
    h_levels = _crs_levels(engine, t, True)
    vk = ver.setup_vk(t)
    z = o.rand_scalars(1 << nv, 1910 + nv)
    r = o.rand_scalars(nv, 1920 + nv)
    poly = sqrt_pst.Polynomial.from_evaluations(h.scalars_to_np(z, mont=True))
    v = poly.eval(r)
    ck = sqrt_pst.CommitterKey.from_points(g_levels[0]).with_levels(g_levels, h_levels)
    comm_list, t_gt = poly.commit(ck)
    assert t_gt is not None
    # t = prod e(C_i, h_i) against the oracle's pairing
    h_vec = [o2.affine_from_words(w) for w in h_levels[odd]]
    T = pr.from_words(t_gt)
    assert T == pr.multi_pairing([h.pt_from_np(c) for c in comm_list], h_vec)
    opened = poly.open(shared_transcript(), comm_list, ck, r, t_gt)
    mp = opened.mipp
    assert len(mp.comms_t) == m_col and len(mp.comms_u) == m_col

    def proof_dict():
        return {
            "comms_u": [(h.pt_from_np(l), h.pt_from_np(rr)) for l, rr in mp.comms_u],
            "comms_t": [(pr.from_words(l), pr.from_words(rr)) for l, rr in mp.comms_t],
            "final_a": h.pt_from_np(mp.final_a),
            "final_h": o2.affine_from_words(mp.final_h),
            "pst_proof_h": [h.pt_from_np(p) for p in mp.pst_proof_h],
        }

    U = h.pt_from_np(opened.u)
    pst_proof = [o2.affine_from_words(p) for p in opened.pst_proof]
    assert ver.sqrt_pst_verify(vk, shared_transcript(), U, r, v, pst_proof, proof_dict(), T) is True
    # negative cases: wrong value, tampered PST proof, tampered cross pairing product, wrong T
    assert ver.sqrt_pst_verify(vk, shared_transcript(), U, r, (v + 1) % o.R_ORDER, pst_proof, proof_dict(), T) is False
    bad = list(pst_proof)
    bad[0] = o2.add(bad[0], o2.G2)
    assert ver.sqrt_pst_verify(vk, shared_transcript(), U, r, v, bad, proof_dict(), T) is False
    pd = proof_dict()
    pd["comms_t"][0] = (pr.f12_mul(pd["comms_t"][0][0], T), pd["comms_t"][0][1])
    assert ver.sqrt_pst_verify(vk, shared_transcript(), U, r, v, pst_proof, pd, T) is False
    assert ver.sqrt_pst_verify(vk, shared_transcript(), U, r, v, pst_proof, proof_dict(), pr.f12_sqr(T)) is False
    ck.close()


@pytest.mark.parametrize("nv", [4, 5])
def test_commit_open_bit_exact_vs_oracle_prover(engine, nv):
    """Every value `Polynomial::commit` and `Polynomial::open` produce (comm_list, t, U, the G2 PST proof, comms_u,
    comms_t, final_a, final_h, pst_proof_h) from the GPU path, compared limb for limb with the big-integer restatement of
    the reference's prover (oracle/sqrt_pst.py) on the same CRS, polynomial, point and transcript."""
    from oracle import bls12_377_g2 as o2
    from oracle import pairing as pr
    from oracle import sqrt_pst as osp

    m_col = nv // 2
    m_row = nv - m_col
    t = o.rand_scalars(m_row, 3900 + nv)
    ock = osp.setup_ck(t)
    g_levels = [h.pts_to_np(l) for l in ock["powers_of_g"]]
    h_levels = [np.array([o2.affine_to_words(p) for p in l], dtype=np.uint64).reshape(-1, 24) for l in ock["powers_of_h"]]
    z = o.rand_scalars(1 << nv, 3910 + nv)
    r = o.rand_scalars(nv, 3920 + nv)
    # oracle
    opoly = osp.Polynomial(z)
    o_comm, o_T = opoly.commit(ock)
    o_U, o_pst, o_mipp = opoly.open(shared_transcript(), o_comm, ock, r, o_T)
    # GPU
    poly = sqrt_pst.Polynomial.from_evaluations(h.scalars_to_np(z, mont=True))
    ck = sqrt_pst.CommitterKey.from_points(g_levels[0]).with_levels(g_levels, h_levels)
    comm_list, t_gt = poly.commit(ck)
    opened = poly.open(shared_transcript(), comm_list, ck, r, t_gt)
    assert np.array_equal(comm_list, h.pts_to_np(o_comm))
    assert np.array_equal(t_gt, np.array(pr.to_words(o_T), dtype=np.uint64))
    assert np.array_equal(opened.u, h.pts_to_np([o_U])[0])
    assert np.array_equal(opened.pst_proof, np.array([o2.affine_to_words(p) for p in o_pst], dtype=np.uint64))
    mp = opened.mipp
    assert mp.xs_inv == o_mipp["xs_inv"] and mp.rs == o_mipp["rs"]
    for (gl, gr), (ol, orr) in zip(mp.comms_u, o_mipp["comms_u"]):
        assert np.array_equal(gl, h.pts_to_np([ol])[0]) and np.array_equal(gr, h.pts_to_np([orr])[0])
    for (gl, gr), (ol, orr) in zip(mp.comms_t, o_mipp["comms_t"]):
        assert np.array_equal(gl, np.array(pr.to_words(ol), dtype=np.uint64))
        assert np.array_equal(gr, np.array(pr.to_words(orr), dtype=np.uint64))
    assert len(mp.comms_t) == len(o_mipp["comms_t"]) == m_col
    assert np.array_equal(mp.final_a, h.pts_to_np([o_mipp["final_a"]])[0])
    assert np.array_equal(mp.final_h, np.array(o2.affine_to_words(o_mipp["final_h"]), dtype=np.uint64))
    assert np.array_equal(mp.pst_proof_h, h.pts_to_np(o_mipp["pst_proof_h"]))
    ck.close()


@pytest.mark.parametrize("nv", [5, 6])
def test_roundtrip_with_the_poseidon_transcript(engine, nv):
    """check_sqrt_poly_commit (src/sqrt_pst.rs:297-342) with the reference's OWN transcript on both sides: the prover
    mirror draws its challenges from `PoseidonTranscript<Fq>` (C++ sponge, serialize.py encodings), the oracle verifier
    from the independent Python sponge and its own encodings (oracle/poseidon.py). It only verifies if both sides
    absorbed identical bytes and every prover value is right."""
    from oracle import bls12_377_g2 as o2
    from oracle import pairing as pr
    from oracle import poseidon as op
    from oracle import verifier as ver
    from testudo_b200 import poseidon_transcript as pt

    m_col = nv // 2
    m_row = nv - m_col
    t = o.rand_scalars(m_row, 2900 + nv)
    g_levels = _crs_levels(engine, t, False)
    h_levels = _crs_levels(engine, t, True)
    vk = ver.setup_vk(t)
    z = o.rand_scalars(1 << nv, 2910 + nv)
    r = o.rand_scalars(nv, 2920 + nv)
    poly = sqrt_pst.Polynomial.from_evaluations(h.scalars_to_np(z, mont=True))
    v = poly.eval(r)
    ck = sqrt_pst.CommitterKey.from_points(g_levels[0]).with_levels(g_levels, h_levels)
    comm_list, t_gt = poly.commit(ck)
    prover = pt.PoseidonTranscript("fq")
    opened = poly.open(prover.as_challenge(), comm_list, ck, r, t_gt)
    mp = opened.mipp
    proof = {
        "comms_u": [(h.pt_from_np(l), h.pt_from_np(rr)) for l, rr in mp.comms_u],
        "comms_t": [(pr.from_words(l), pr.from_words(rr)) for l, rr in mp.comms_t],
        "final_a": h.pt_from_np(mp.final_a),
        "final_h": o2.affine_from_words(mp.final_h),
        "pst_proof_h": [h.pt_from_np(p) for p in mp.pst_proof_h],
    }
    U = h.pt_from_np(opened.u)
    pst_proof = [o2.affine_from_words(p) for p in opened.pst_proof]
    T = pr.from_words(t_gt)
    ark, mds = pt.reference_parameters()
    assert ver.sqrt_pst_verify(vk, op.OracleTranscript(ark, mds).challenge, U, r, v, pst_proof, proof, T) is True
    assert all(0 < c < o.R_ORDER for c in mp.xs_inv) and len(set(mp.xs_inv)) == len(mp.xs_inv)
    bad = dict(proof)
    bad["comms_u"] = [(proof["comms_u"][0][1], proof["comms_u"][0][0])] + proof["comms_u"][1:]
    assert ver.sqrt_pst_verify(vk, op.OracleTranscript(ark, mds).challenge, U, r, v, pst_proof, bad, T) is False
    ck.close()
