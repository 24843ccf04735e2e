"""GPU parity at the sizes BASELINE.json quotes (configs[1..3]), against the ORACLE -- not against the engine's own scalar
multiplication: the single MSM at 2^24 points, `Polynomial::commit` at nv = 20 (in full) and nv = 26 (sampled rows, an
all-zero row, and t), the MIPP prover at 2^13 commitments with every value checked through closed forms over known
discrete logs (oracle/closed_forms.py), and the reference's round trip commit -> open -> verify (src/sqrt_pst.rs:297-342)
at 2^26 coefficients with the oracle verifier. The engine generates the big synthetic inputs (outer sums of generator
multiples, CRS levels); every test spot-checks those inputs against the oracle before relying on their discrete logs.
"""
import ctypes
import hashlib

import numpy as np
import pytest

import helpers as h
from oracle import bls12_377 as o
from oracle import closed_forms as cf
from testudo_b200 import _lib, curve, fr, mipp, msm, sqrt_pst, synthetic

pytestmark = pytest.mark.gpu
R = o.R_ORDER


def P(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def oracle_multiples(oracle_c):
    """start, step, count -> [(start + i step) G] by the C oracle's point additions (one big-int mul each for the ends)."""
    def gen(start, step, count):
        s = h.pts_to_np([o.mul(start % R, o.G)])[0]
        d = h.pts_to_np([o.mul(step % R, o.G)])[0]
        return oracle_c.gen_points(s, d, count)
    return gen


def base_dlog(seed, n, idx):
    a0, sa, b0, sb = synthetic.dlog_params(seed)
    _, nb = synthetic.split(n)
    return (a0 + (idx // nb) * sa + b0 + (idx % nb) * sb) % R


def spot_check_bases(bases, n, seed, count=8):
    rng = np.random.default_rng(seed + 17)
    for idx in [0, 1, n - 1] + [int(x) for x in rng.integers(0, n, size=count)]:
        assert h.pt_from_np(bases[idx].cpu().numpy().view(np.uint64)) == o.mul(base_dlog(seed, n, idx), o.G), idx


def transcript(tag=b"baseline-sizes"):
    state = hashlib.sha256(tag)
    seen = []

    def challenge(label, values):
        state.update(label)
        for v in values:
            state.update(np.asarray(v, dtype=np.uint64).tobytes())
        c = int.from_bytes(state.digest(), "little") % R or 1
        seen.append((label, c))
        return c

    challenge.seen = seen
    return challenge


# ---- configs[1]: single G1 MSM, 2^24 points ---------------------------------------------------------------------------------
@pytest.mark.parametrize("skew", [False, True], ids=["uniform", "skewed_50_25_25"])
def test_msm_2p24_vs_oracle(engine, oracle_c, skew):
    import torch

    n = 1 << 24
    seed = 2401 + int(skew)
    bases = synthetic.make_bases_dev(n, seed=seed, multiples_of_g=oracle_multiples(oracle_c))
    spot_check_bases(bases, n, seed)
    sc = synthetic.make_scalars_dev(n, seed=seed + 100, skew=skew)
    want = o.mul(synthetic.expected_dlog(sc, n, seed=seed), o.G)
    out = torch.zeros(12, dtype=torch.int64, device="cuda")
    _lib.check(engine.tb200_msm_g1_dev(ctypes.c_void_p(bases.data_ptr()), ctypes.c_void_p(sc.data_ptr()), n, 0,
                                       ctypes.c_void_p(out.data_ptr()), None))
    _lib.check(engine.tb200_stream_sync())
    assert h.pt_from_np(out.cpu().numpy().view(np.uint64)) == want
    if not skew:
        # the host-facing path at full size: chunked upload overlapped with the sort / accumulate of the previous chunk
        got = msm.msm_bigint(bases.cpu().numpy().view(np.uint64), sc.cpu().numpy().view(np.uint64))
        assert h.pt_from_np(got) == want
    del bases, sc
    torch.cuda.empty_cache()


# ---- configs[0]: Polynomial::commit at 2^20 coefficients, every row against the C oracle ----------------------------------------
def test_commit_nv20_every_row_vs_c_oracle(engine, oracle_c):
    nv = 20
    rows = cols = 1 << 10
    srs = oracle_c.gen_points(h.pts_to_np([o.mul(0xABCDEF, o.G)])[0], h.pts_to_np([o.mul(0x1357, o.G)])[0], cols)
    z = h.np_rand_scalars(1 << nv, 2020)                      # canonical values < r are valid Montgomery limbs too
    z[7::rows] = 0                                            # row 7 is the zero polynomial
    z[(np.arange(cols) << 10) | 9] = np.array([1, 0, 0, 0], dtype=np.uint64)
    ck = sqrt_pst.CommitterKey.from_points(srs)
    poly = sqrt_pst.Polynomial.from_evaluations(z)
    comm_list, _ = poly.commit(ck)
    exp = oracle_c.msm_g1_batch(srs, z, rows, cols, 1, rows, mont=True)
    assert np.array_equal(comm_list, exp)
    assert not comm_list[7].any()
    ck.close()


# ---- configs[2]: Polynomial::commit at 2^26 coefficients ------------------------------------------------------------------------
def test_commit_nv26_sampled_rows_zero_row_and_t(engine, oracle_c):
    import torch

    from oracle import bls12_377_g2 as o2
    from oracle import pairing as pr

    nv, rows, cols = 26, 1 << 13, 1 << 13
    srs = oracle_c.gen_points(h.pts_to_np([o.mul(0x26262626, o.G)])[0], h.pts_to_np([o.mul(0x777, o.G)])[0], cols)
    zd = synthetic.make_scalars_dev(1 << nv, seed=2626)
    zd.view(cols, rows, 4)[:, 11] = 0                         # row 11 (Z[(j << m_col) | 11]) is the zero polynomial
    z = zd.cpu().numpy().view(np.uint64)
    del zd
    torch.cuda.empty_cache()
    ks = [(0x9E3779B97F4A7C15 * (i + 1)) % R for i in range(32)]
    h32 = np.array([o2.affine_to_words(o2.mul(k, o2.G2)) for k in ks], dtype=np.uint64)
    h_vec = np.ascontiguousarray(np.tile(h32, (rows // 32, 1)))
    ck = sqrt_pst.CommitterKey.from_points(srs)
    ck.powers_of_h = [h_vec]                                  # commit only reads powers_of_h[odd] (src/sqrt_pst.rs:128)
    poly = sqrt_pst.Polynomial.from_evaluations(z)
    comm_list, t_gt = poly.commit(ck)
    sample = sorted({0, 1, 11, rows - 1} | {int(x) for x in np.random.default_rng(26).integers(0, rows, size=64)})
    zs = np.ascontiguousarray(np.stack([z[i::rows] for i in sample]))          # [len(sample), cols, 4]
    exp = oracle_c.msm_g1_batch(srs, zs.reshape(-1, 4), len(sample), cols, cols, 1, mont=True)
    assert np.array_equal(comm_list[sample], exp)
    assert not comm_list[11].any()
    # t = prod_i e(C_i, k_i G2) = e(sum_i k_i C_i, G2): the fold by the C oracle's MSM, the pairing by the oracle
    folded = oracle_c.msm_g1(comm_list, h.scalars_to_np([ks[i % 32] for i in range(rows)]))
    assert pr.from_words(t_gt) == pr.pairing(h.pt_from_np(folded), o2.G2)
    ck.close()


# ---- configs[2]/[3]: the MIPP prover over 2^13 commitments, every value through closed forms ------------------------------------
def _g2_multiples(engine, dlogs):
    from oracle import bls12_377_g2 as o2

    gen = np.ascontiguousarray(np.tile(np.array([o2.affine_to_words(o2.G2)], dtype=np.uint64), (len(dlogs), 1)))
    out = np.zeros((len(dlogs), 24), dtype=np.uint64)
    _lib.check(engine.tb200_test_g2_mul(P(gen), P(h.scalars_to_np(dlogs)), len(dlogs), P(out)))
    for idx in (0, 1, len(dlogs) // 2, len(dlogs) - 1):       # generated by the engine: spot-check against the oracle
        assert o2.affine_from_words(out[idx]) == o2.mul(dlogs[idx], o2.G2)
    return out


def test_mipp_2p13_every_value_vs_closed_forms(engine, oracle_c):
    from oracle import bls12_377_g2 as o2
    from oracle import pairing as pr

    n = 1 << 13
    seed = 1313
    a_dev = synthetic.make_bases_dev(n, seed=seed, multiples_of_g=oracle_multiples(oracle_c))
    spot_check_bases(a_dev, n, seed, count=4)
    a = a_dev.cpu().numpy().view(np.uint64)
    alphas = [base_dlog(seed, n, i) for i in range(n)]
    etas = o.rand_scalars(n, seed + 1)
    hk = _g2_multiples(engine, etas)
    ys = o.rand_scalars(n, seed + 2)
    U = h.pts_to_np([o.mul(sum(x * y for x, y in zip(alphas, ys)) % R, o.G)])[0]
    ch = transcript(b"mipp-2p13")
    proof = mipp.MippProofG1.prove(ch, a, h.scalars_to_np(ys, mont=True), U, h=hk)
    c_invs = [c for label, c in ch.seen if label == b"challenge_i"]
    assert len(c_invs) == 13 and proof.xs_inv == c_invs
    want = cf.mipp_dlogs(alphas, etas, ys, c_invs)
    for k in range(13):
        assert h.pt_from_np(proof.comms_u[k][0]) == cf.g1_of(want["u"][k][0]), f"comm_u_l round {k}"
        assert h.pt_from_np(proof.comms_u[k][1]) == cf.g1_of(want["u"][k][1]), f"comm_u_r round {k}"
        assert pr.from_words(proof.comms_t[k][0]) == cf.gt_of(want["t"][k][0]), f"comm_t_l round {k}"
        assert pr.from_words(proof.comms_t[k][1]) == cf.gt_of(want["t"][k][1]), f"comm_t_r round {k}"
    assert h.pt_from_np(proof.final_a) == cf.g1_of(want["final_a"])
    assert o2.affine_from_words(proof.final_h) == cf.g2_of(want["final_h"])
    assert fr.from_mont_words(proof.final_y.reshape(1, 4))[0] == want["final_y"]


# ---- configs[2]: the reference's round trip at 2^26 coefficients, verified by the oracle verifier -------------------------------
def _crs_levels(engine, t, g2):
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
        for idx in (0, len(e) - 1):                             # engine-generated CRS: spot-check against the oracle
            if g2:
                assert o2.affine_from_words(out[idx]) == o2.mul(e[idx], o2.G2)
            else:
                assert h.pt_from_np(out[idx]) == o.mul(e[idx], o.G)
        levels.append(out)
    return levels


def _neutral(values):
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


def test_sqrt_pst_nv26_commit_open_accepted_by_oracle_verifier(engine):
    """check_sqrt_poly_commit (src/sqrt_pst.rs:297-342) at BASELINE configs[2]'s size: every prover value from the GPU
    (k_miller's thread-per-pair path and the endomorphism folds at production size), the verifier = oracle/verifier.py."""
    import torch

    from oracle import bls12_377_g2 as o2
    from oracle import pairing as pr
    from oracle import pst
    from oracle import verifier as ver

    nv = 26
    m_col = m_row = 13
    t = o.rand_scalars(m_row, 2600)
    g_levels = _crs_levels(engine, t, False)
    h_levels = _crs_levels(engine, t, True)
    vk = ver.setup_vk(t)
    zd = synthetic.make_scalars_dev(1 << nv, seed=2601)
    z = zd.cpu().numpy().view(np.uint64)
    del zd
    torch.cuda.empty_cache()
    r = o.rand_scalars(nv, 2602)
    poly = sqrt_pst.Polynomial.from_evaluations(z)
    v = poly.eval(r)
    ck = sqrt_pst.CommitterKey.from_points(g_levels[0]).with_levels(g_levels, h_levels)
    comm_list, t_gt = poly.commit(ck)
    inner = transcript(b"nv26-roundtrip")
    opened = poly.open(inner, comm_list, ck, r, t_gt)
    mp = opened.mipp
    assert len(mp.comms_t) == m_col and len(mp.comms_u) == m_col and len(opened.pst_proof) == m_row
    # closed forms of the two values whose discrete logs need no pass over Z on the host
    q = fr.from_mont_words(poly.q)
    assert h.pt_from_np(opened.u) == o.mul(pst.mle_eval(q, t), o.G)
    ph = mipp.polynomial_evaluations_from_transcript(mp.xs_inv)
    assert o2.affine_from_words(mp.final_h) == o2.mul(pst.mle_eval(ph, t), o2.G2)
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

    def verifier_transcript():
        tr = transcript(b"nv26-roundtrip")
        return lambda label, values: tr(label, _neutral(values))

    assert ver.sqrt_pst_verify(vk, verifier_transcript(), U, r, v, pst_proof, proof, T) is True
    assert ver.sqrt_pst_verify(vk, verifier_transcript(), U, r, (v + 1) % R, pst_proof, proof, T) is False
    # the engine's own `Polynomial::verify` (src/sqrt_pst.rs:232-267) on the same proof at this size: 27 GT powers, the
    # 27-point UC fold, five pairing products of up to 14 pairs -- the same verdicts as the oracle verifier
    from testudo_b200 import multilinear_pc
    vk_e = multilinear_pc.VerifierKey(
        nv=m_row, g=h.pts_to_np([vk["g"]])[0], h=np.array(o2.affine_to_words(vk["h"]), dtype=np.uint64),
        g_mask_random=h.pts_to_np(vk["g_mask"]),
        h_mask_random=np.array([o2.affine_to_words(p) for p in vk["h_mask"]], dtype=np.uint64))
    same = transcript(b"nv26-roundtrip")
    assert sqrt_pst.Polynomial.verify(same, vk_e, opened.u, r, v, opened.pst_proof, mp, t_gt) is True
    same = transcript(b"nv26-roundtrip")
    assert sqrt_pst.Polynomial.verify(same, vk_e, opened.u, r, (v + 1) % R, opened.pst_proof, mp, t_gt) is False
    ck.close()


def test_msm_overlap_option_gives_the_same_point(engine, oracle_c):
    """tb200_set_msm_overlap(1): window ranges on three streams (resident inputs) and chunk sorts on a side stream (host
    inputs) -- off by default (no measurable gain), but the results must be the points the default path gives."""
    import torch

    n = 1 << 21
    seed = 2121
    bases = synthetic.make_bases_dev(n, seed=seed, multiples_of_g=oracle_multiples(oracle_c))
    sc = synthetic.make_scalars_dev(n, seed=seed + 1, skew=True)
    want = o.mul(synthetic.expected_dlog(sc, n, seed=seed), o.G)
    out = torch.zeros(12, dtype=torch.int64, device="cuda")
    engine.tb200_set_msm_overlap(1)
    try:
        for _ in range(2):                                    # twice: the side arenas and events are reused
            _lib.check(engine.tb200_msm_g1_dev(ctypes.c_void_p(bases.data_ptr()), ctypes.c_void_p(sc.data_ptr()), n, 0,
                                               ctypes.c_void_p(out.data_ptr()), None))
            _lib.check(engine.tb200_stream_sync())
            assert h.pt_from_np(out.cpu().numpy().view(np.uint64)) == want
            got = msm.msm_bigint(bases.cpu().numpy().view(np.uint64), sc.cpu().numpy().view(np.uint64))
            assert h.pt_from_np(got) == want
    finally:
        engine.tb200_set_msm_overlap(0)
    del bases, sc
    torch.cuda.empty_cache()
