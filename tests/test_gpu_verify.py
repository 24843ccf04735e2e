"""GPU parity of the VERIFIER side of the commitment path: `Polynomial::verify` (src/sqrt_pst.rs:232-267) ->
`MippProof::verify` (src/mipp.rs:182-333) + `MultilinearPC::check` / the fork's `check_2`, and the two entry points
added for it (tb200_gt_multi_pow, tb200_msm_g1_each). The checker is oracle/verifier.py (big integers, its own sponge and
encodings): both verifiers must return the same answer on the proofs the GPU prover makes -- honest and tampered -- and
the new primitives must equal the oracle's values limb for limb. The reference's own round trip is
`check_sqrt_poly_commit` (src/sqrt_pst.rs:297-342) and benches/pst.rs:76-90."""
import ctypes

import numpy as np
import pytest

import helpers as h
from oracle import bls12_377 as o
from oracle import bls12_377_g2 as o2
from oracle import pairing as pr
from oracle import poseidon as op
from oracle import verifier as ver
from testudo_b200 import _lib, msm, multilinear_pc, pairing, sqrt_pst
from testudo_b200 import poseidon_transcript as pt

pytestmark = pytest.mark.gpu


def P(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def g2_np(points) -> np.ndarray:
    return np.array([o2.affine_to_words(p) for p in points], dtype=np.uint64).reshape(-1, 24)


def gt_np(x) -> np.ndarray:
    return np.array(pr.to_words(x), dtype=np.uint64)


def vk_np(vk) -> multilinear_pc.VerifierKey:
    return multilinear_pc.VerifierKey(nv=vk["nv"], g=h.pts_to_np([vk["g"]])[0], h=g2_np([vk["h"]])[0],
                                      g_mask_random=h.pts_to_np(vk["g_mask"]), h_mask_random=g2_np(vk["h_mask"]))


def crs_levels(engine, t, g2):
    """powers[k][x] = eq((t_k..t_{nv-1}), x) * generator (ark-poly-commit `setup`), through the engine's test multiplier"""
    gen = g2_np([o2.G2]) if g2 else h.pts_to_np([o.G])
    fn = engine.tb200_test_g2_mul if g2 else engine.tb200_test_g1_mul
    out = []
    for k in range(len(t)):
        e = [1]
        for tj in t[k:]:
            e = [v * (1 - tj) % o.R_ORDER for v in e] + [v * tj % o.R_ORDER for v in e]
        pts = np.zeros((len(e), gen.shape[1]), dtype=np.uint64)
        g = np.ascontiguousarray(np.tile(gen, (len(e), 1)))
        _lib.check(fn(P(g), P(h.scalars_to_np(e)), len(e), P(pts)))
        out.append(pts)
    return out


def test_msm_g1_each_against_the_oracle(engine):
    """rows of 1, 2, 3 and 8 points: identity bases, zero scalars, r - 1, P - P, and random rows"""
    pts, _ = o.rand_points(24, 771)
    for per_row in (1, 2, 3, 8):
        rows = 24 // per_row
        ks = o.rand_scalars(24, 772 + per_row)
        bases = list(pts)
        ks[0] = 0
        ks[1] = o.R_ORDER - 1
        bases[2] = None                                    # the identity as a base: the all-zero row
        if per_row >= 2:                                   # last row: k P + (r - k) P = identity
            bases[-1] = bases[-2]
            ks[-1] = (o.R_ORDER - ks[-2]) % o.R_ORDER
        got = msm.msm_each(h.pts_to_np(bases), h.scalars_to_np(ks), per_row)
        assert got.shape == (rows, 12)
        for i in range(rows):
            want = o.msm_naive(bases[i * per_row:(i + 1) * per_row], ks[i * per_row:(i + 1) * per_row])
            assert h.pt_from_np(got[i]) == want, (per_row, i)
        if per_row == 2:
            assert not got[-1].any()                       # the identity is the all-zero row
    with pytest.raises(ValueError):
        msm.msm_each(h.pts_to_np(pts[:3]), h.scalars_to_np([1, 2, 3]), 2)
    assert engine.tb200_msm_g1_each(P(h.pts_to_np(pts[:9])), P(h.scalars_to_np(list(range(9)))), 1, 9, 0,
                                    P(np.zeros(12, dtype=np.uint64))) == -1
    assert engine.tb200_msm_g1_each(None, None, 0, 2, 0, None) == 0


def test_gt_multi_pow_against_the_oracle(engine):
    """prod_i base_i^e_i vs the oracle tower: unitary values (pairing outputs) and ARBITRARY field elements (the kernel must
    not rely on cyclotomic squarings), exponents 0, 1, r - 1, random; canonical and Montgomery exponents; and the
    single-element tb200_gt_pow now on the same cooperative kernel."""
    import random
    rng = random.Random(99)
    e_gen = pr.pairing(o.G, o2.G2)
    bases = [pr.f12_pow(e_gen, rng.randrange(1, o.R_ORDER)) for _ in range(3)]
    bases += [tuple((rng.randrange(o.Q), rng.randrange(o.Q)) for _ in range(6)) for _ in range(3)]
    exps = [0, 1, o.R_ORDER - 1, rng.randrange(o.R_ORDER), rng.randrange(o.R_ORDER), 2]
    want = pr.F12_ONE
    each = []
    for b, e in zip(bases, exps):
        each.append(pr.f12_pow(b, e))
        want = pr.f12_mul(want, each[-1])
    B = np.stack([gt_np(b) for b in bases])
    got = pairing.gt_multi_pow(B, h.scalars_to_np(exps))
    assert pr.from_words(got) == want
    got_m = pairing.gt_multi_pow(B, h.scalars_to_np(exps, mont=True), mont=True)
    assert np.array_equal(got, got_m)
    pw = pairing.gt_pow(B, h.scalars_to_np(exps))
    assert [pr.from_words(r) for r in pw] == each
    one = pairing.gt_multi_pow(np.zeros((0, 72), dtype=np.uint64), np.zeros((0, 4), dtype=np.uint64))
    assert pr.from_words(one) == pr.F12_ONE
    assert pr.from_words(pairing.gt_multi_pow(B[:1], h.scalars_to_np([7]))) == pr.f12_pow(bases[0], 7)


def test_multi_pairing_batch_against_the_oracle(engine):
    """products of 1, 3 and 0 pairs (padded with identity pairs) and one holding an identity on either side, in one pass,
    vs the oracle's multi_pairing and vs the single-product entry point"""
    ps, _ = o.rand_points(5, 881)
    qs = [o2.mul(k, o2.G2) for k in o.rand_scalars(5, 882)]
    products = [(ps[:1], qs[:1]), (ps[1:4], qs[1:4]), ([], []), ([ps[4], None, ps[0]], [qs[4], qs[1], None])]
    got = pairing.multi_pairing_batch([(h.pts_to_np(p), g2_np(q)) for p, q in products])
    assert got.shape == (4, 72)
    for row, (p, q) in zip(got, products):
        assert pr.from_words(row) == pr.multi_pairing(list(p), list(q))
        assert np.array_equal(row, pairing.multi_pairing(h.pts_to_np(p), g2_np(q)))
    assert pairing.multi_pairing_batch([]).shape == (0, 72)
    assert engine.tb200_multi_pairing_batch(None, None, 2, 3, P(np.zeros(144, dtype=np.uint64))) == -1


def _prove(engine, nv, seed):
    m_col = nv // 2
    m_row = nv - m_col
    t = o.rand_scalars(m_row, seed)
    g_levels = crs_levels(engine, t, False)
    h_levels = crs_levels(engine, t, True)
    vk = ver.setup_vk(t)
    z = o.rand_scalars(1 << nv, seed + 10)
    r = o.rand_scalars(nv, seed + 20)
    poly = sqrt_pst.Polynomial.from_evaluations(h.scalars_to_np(z, mont=True))
    v = poly.eval(r)
    ck = sqrt_pst.CommitterKey.from_points(g_levels[0]).with_levels(g_levels, h_levels)
    comm_list, t_gt = poly.commit(ck)
    prover = pt.PoseidonTranscript("fq")
    opened = poly.open(prover.as_challenge(), comm_list, ck, r, t_gt)
    ck.close()
    return vk, r, v, opened, t_gt


def _oracle_verdict(vk, r, v, opened, t_gt) -> bool:
    mp = opened.mipp
    proof = {
        "comms_u": [(h.pt_from_np(l), h.pt_from_np(rr)) for l, rr in mp.comms_u],
        "comms_t": [(pr.from_words(l), pr.from_words(rr)) for l, rr in mp.comms_t],
        "final_a": h.pt_from_np(mp.final_a),
        "final_h": o2.affine_from_words(mp.final_h),
        "pst_proof_h": [h.pt_from_np(p) for p in mp.pst_proof_h],
    }
    ark, mds = pt.reference_parameters()
    return ver.sqrt_pst_verify(vk, op.OracleTranscript(ark, mds).challenge, h.pt_from_np(opened.u), r, v,
                               [o2.affine_from_words(p) for p in opened.pst_proof], proof, pr.from_words(t_gt))


def _engine_verdict(vk, r, v, opened, t_gt) -> bool:
    return sqrt_pst.Polynomial.verify(pt.PoseidonTranscript("fq").as_challenge(), vk_np(vk), opened.u, r, v,
                                      opened.pst_proof, opened.mipp, t_gt)


@pytest.mark.parametrize("nv", [4, 5, 6])
def test_engine_verifier_agrees_with_the_oracle_verifier(engine, nv):
    """commit -> open -> verify == true (check_sqrt_poly_commit, src/sqrt_pst.rs:297-342; benches/pst.rs:76-90) with the
    reference's Poseidon transcript on both sides, then one field of the proof at a time is damaged: the engine's verifier
    and the oracle's must give the same verdict every time, and reject every damaged proof."""
    import copy
    vk, r, v, opened, t_gt = _prove(engine, nv, 3100 + 37 * nv)
    assert _engine_verdict(vk, r, v, opened, t_gt) is True
    assert _oracle_verdict(vk, r, v, opened, t_gt) is True

    def damaged():
        d = copy.deepcopy(opened)
        yield "value", d, (v + 1) % o.R_ORDER, r, t_gt
        d = copy.deepcopy(opened)
        d.mipp.comms_u[0] = (d.mipp.comms_u[0][1], d.mipp.comms_u[0][0])
        yield "comm_u swapped", d, v, r, t_gt
        d = copy.deepcopy(opened)
        d.mipp.comms_t[-1] = (d.mipp.comms_t[-1][1], d.mipp.comms_t[-1][0])
        yield "comm_t swapped", d, v, r, t_gt
        d = copy.deepcopy(opened)
        d.pst_proof = d.pst_proof[::-1].copy()
        yield "pst_proof reversed", d, v, r, t_gt
        d = copy.deepcopy(opened)
        d.mipp.final_a = h.pts_to_np([o.mul(2, h.pt_from_np(d.mipp.final_a))])[0]
        yield "final_a doubled", d, v, r, t_gt
        d = copy.deepcopy(opened)
        d.mipp.final_h = g2_np([o2.mul(3, o2.affine_from_words(d.mipp.final_h))])[0]
        yield "final_h tripled", d, v, r, t_gt
        d = copy.deepcopy(opened)
        d.mipp.pst_proof_h = d.mipp.pst_proof_h[::-1].copy()
        yield "pst_proof_h reversed", d, v, r, t_gt
        d = copy.deepcopy(opened)
        d.u = h.pts_to_np([o.mul(5, h.pt_from_np(d.u))])[0]
        yield "U", d, v, r, t_gt
        yield "T", copy.deepcopy(opened), v, r, gt_np(pr.f12_sqr(pr.from_words(t_gt)))
        yield "point", copy.deepcopy(opened), v, [(r[0] + 1) % o.R_ORDER] + list(r[1:]), t_gt

    # over the wire: compressed `Proof` / `MippProof` / `Commitment` bytes (benches/pst.rs:64-74) -> deserialised -> verified
    from testudo_b200 import mipp, serialize
    wire_mipp = mipp.MippProofG1.from_bytes(opened.mipp.to_bytes())
    wire_pst = serialize.pst_proof_from_bytes(serialize.pst_proof_bytes(opened.pst_proof))
    _, wire_u = serialize.commitment_from_bytes(serialize.commitment_bytes(nv - nv // 2, opened.u))
    assert sqrt_pst.Polynomial.verify(pt.PoseidonTranscript("fq").as_challenge(), vk_np(vk), wire_u, r, v, wire_pst,
                                      wire_mipp, t_gt) is True
    b = list(r[nv // 2 + nv % 2:])
    assert opened.mipp.verify(vk_np(vk), pt.PoseidonTranscript("fq").as_challenge(), b, opened.u, t_gt) is True   # src/mipp.rs:182
    for what, d, vv, rr, tt in damaged():
        got = _engine_verdict(vk, rr, vv, d, tt)
        assert got is False, what
        if what in ("value", "comm_u swapped", "pst_proof reversed", "final_a doubled", "T"):   # the oracle costs seconds each
            assert _oracle_verdict(vk, rr, vv, d, tt) is False, what


def test_check_and_check_2_values_equal_the_oracle_s(engine):
    """Both sides of the two pairing equations as GT ELEMENTS against oracle/verifier.py's restatement (the G1-side fold of
    check_2 must give the very field element the reference's G2-side combination gives), on a claim that holds and on one
    that does not."""
    nv = 3
    t = o.rand_scalars(nv, 4200)
    vk = ver.setup_vk(t)
    vkn = vk_np(vk)
    z = o.rand_scalars(nv, 4201)
    # check: any commitment / proof values exercise the equation's two sides
    C = o.mul(o.rand_scalars(1, 4202)[0], o.G)
    proofs = [o2.mul(k, o2.G2) for k in o.rand_scalars(nv, 4203)]
    assert multilinear_pc.check(vkn, h.pts_to_np([C])[0], z, 77, g2_np(proofs)) == ver.check(vk, C, z, 77, proofs)
    # an honest opening of the constant polynomial f = 5: C = 5 g, every quotient is zero -> proofs are the identity
    C5 = o.mul(5, o.G)
    zero_proofs = np.zeros((nv, 24), dtype=np.uint64)
    assert multilinear_pc.check(vkn, h.pts_to_np([C5])[0], z, 5, zero_proofs) is True
    assert multilinear_pc.check(vkn, h.pts_to_np([C5])[0], z, 6, zero_proofs) is False
    # check_2 with a shorter point (off = 1): f(x) = x_0 over the last two variables of the key, C_h = t_1 h;
    # f(z) - f(t) = (z_0 - t_1) * 1 -> quotient q_0 = 1 (proof g), q_1 = 0 (identity); ark's proof convention is
    # exercised end to end by the round trip above, here only the two sides are compared with the oracle's
    m = 2
    z2 = o.rand_scalars(m, 4204)
    Ch = o2.mul(o.rand_scalars(1, 4205)[0], o2.G2)
    p1 = [o.mul(k, o.G) for k in o.rand_scalars(m, 4206)]
    got = multilinear_pc.check_2(vkn, g2_np([Ch])[0], z2, 91, h.pts_to_np(p1))
    assert got == ver.check_2(vk, Ch, z2, 91, p1)
    # the two sides as values: left and right of the oracle's equation vs the engine's primitives
    off = nv - m
    want_right = pr.multi_pairing(p1, [o2.add(vk["h_mask"][off + i], o2.neg(o2.mul(z2[i], vk["h"]))) for i in range(m)])
    folded = msm.msm_bigint(h.pts_to_np(p1), h.scalars_to_np([(-x) % o.R_ORDER for x in z2]))
    got_right = pairing.multi_pairing(np.concatenate([h.pts_to_np(p1), folded.reshape(1, 12)]),
                                      np.concatenate([vkn.h_mask_random[off:], vkn.h.reshape(1, 24)]))
    assert pr.from_words(got_right) == want_right


def test_batched_pairing_and_power_paths_at_larger_sizes(engine):
    """The branches the verifier's own sizes do not reach: batches above 512 pairs (two pairs per warp, even segment
    length; the one-pair-per-warp fallback for an odd length) against the single-product entry point (itself pinned to the
    oracle in tests/test_gpu_pairing.py), and more than 8192 GT powers (thread-per-element kernel) against the same
    product taken in two halves on the cooperative kernel."""
    ps, _ = o.rand_points(8, 991)
    qs = [o2.mul(k, o2.G2) for k in o.rand_scalars(8, 992)]
    P8, Q8 = h.pts_to_np(ps), g2_np(qs)
    for each in (300, 201):
        reps = (each * 3 + 7) // 8
        A = np.tile(P8, (reps, 1))[: each * 3].copy()
        B = np.tile(Q8, (reps, 1))[: each * 3].copy()
        A[5] = 0                                            # an identity inside the first product
        got = pairing.multi_pairing_batch([(A[i * each:(i + 1) * each], B[i * each:(i + 1) * each]) for i in range(3)])
        for i in range(3):
            assert np.array_equal(got[i], pairing.multi_pairing(A[i * each:(i + 1) * each], B[i * each:(i + 1) * each])), (each, i)
    n = 8200
    gts = np.stack([pairing.pairing(P8[i], Q8[i]) for i in range(4)])
    bases = np.tile(gts, (n // 4, 1))
    exps = h.scalars_to_np([(7 * i + 3) % 65521 for i in range(n)])
    whole = pairing.gt_multi_pow(bases, exps)
    lo = pairing.gt_multi_pow(bases[:4100], exps[:4100])
    hi = pairing.gt_multi_pow(bases[4100:], exps[4100:])
    both = pairing.gt_multi_pow(np.stack([lo, hi]), h.scalars_to_np([1, 1]))
    assert np.array_equal(whole, both)
    k = sum((7 * i + 3) % 65521 for i in range(0, n, 4))   # the exponents that meet base 0: one closed form as a spot check
    only0 = pairing.gt_multi_pow(bases[::4], exps[::4])
    assert np.array_equal(only0, pairing.gt_pow(gts[:1], h.scalars_to_np([k % o.R_ORDER]))[0])
