"""CPU test of the HOST LOGIC of the verifier mirror (testudo_b200/{sqrt_pst,mipp,multilinear_pc}.py: transcript replay,
index conventions, the ragged row batch, the G1-side form of check_2, the order of the five pairing products) with the
engine's entry points replaced by big-integer stand-ins built from the oracle. The proof comes from the oracle's prover
(oracle/sqrt_pst.py), the reference verdict from the oracle's verifier (oracle/verifier.py): `Polynomial.verify` must accept
what it accepts and reject what it rejects. No GPU, no library call: the CUDA entry points themselves are pinned to the
oracle by tests/test_gpu_verify.py."""
import hashlib

import numpy as np
import pytest

import helpers as h
from oracle import bls12_377 as o
from oracle import bls12_377_g2 as o2
from oracle import pairing as pr
from oracle import sqrt_pst as osp
from oracle import verifier as ver
from testudo_b200 import mipp, msm, msm_g2, multilinear_pc, pairing, sqrt_pst


def g2_np(points) -> np.ndarray:
    return np.array([o2.affine_to_words(p) for p in points], dtype=np.uint64).reshape(-1, 24)


def gt_np(x) -> np.ndarray:
    return np.array(pr.to_words(x), dtype=np.uint64)


def word_transcript(tag=b"verifier-host-logic"):
    """hashes the C-ABI words of every appended value: the oracle side converts its points, the mirror passes arrays"""
    state = hashlib.sha256(tag)

    def challenge(label, values):
        state.update(label)
        for v in values:
            if isinstance(v, tuple) and v and v[0] in ("g1", "g2", "gt"):
                kind, val = v
                words = o.affine_to_words(val) if kind == "g1" else (o2.affine_to_words(val) if kind == "g2" else pr.to_words(val))
                state.update(np.array(words, dtype=np.uint64).tobytes())
            else:
                state.update(np.ascontiguousarray(v, dtype=np.uint64).tobytes())
        return int.from_bytes(state.digest(), "little") % o.R_ORDER or 1

    return challenge


@pytest.fixture
def stand_ins(monkeypatch):
    """the engine calls the verifier mirror makes, as oracle arithmetic on the same word layouts"""
    calls = {"rows": 0, "pairing_batches": 0, "gt": 0}

    def ints(words):
        return h.scalars_from_np(np.asarray(words, dtype=np.uint64).reshape(-1, 4))

    def msm_rows(bases, bigints, row_lengths):
        calls["rows"] += 1
        b = np.asarray(bases, dtype=np.uint64).reshape(-1, 12)
        s = ints(bigints)
        out, pos = [], 0
        for ln in [int(x) for x in row_lengths]:
            out.append(o.msm_naive([h.pt_from_np(r) for r in b[pos:pos + ln]], s[pos:pos + ln]))
            pos += ln
        return h.pts_to_np(out)

    def g2_msm_bigint(bases, bigints):
        b = np.asarray(bases, dtype=np.uint64).reshape(-1, 24)
        return g2_np([o2.msm_naive([o2.affine_from_words(r) for r in b], ints(bigints))])[0]

    def gt_multi_pow(bases, exps, mont=False):
        assert not mont
        calls["gt"] += 1
        acc = pr.F12_ONE
        for b, e in zip(np.asarray(bases, dtype=np.uint64).reshape(-1, 72), ints(exps)):
            acc = pr.f12_mul(acc, pr.f12_pow(pr.from_words(b), e))
        return gt_np(acc)

    def multi_pairing_batch(products):
        calls["pairing_batches"] += 1
        out = []
        for g1s, g2s in products:
            ps = [h.pt_from_np(r) for r in np.asarray(g1s, dtype=np.uint64).reshape(-1, 12)]
            qs = [o2.affine_from_words(r) for r in np.asarray(g2s, dtype=np.uint64).reshape(-1, 24)]
            n = min(len(ps), len(qs))
            out.append(gt_np(pr.multi_pairing(ps[:n], qs[:n])))
        return np.stack(out) if out else np.zeros((0, 72), dtype=np.uint64)

    monkeypatch.setattr(msm, "msm_rows", msm_rows)
    monkeypatch.setattr(msm_g2, "msm_bigint", g2_msm_bigint)
    monkeypatch.setattr(pairing, "gt_multi_pow", gt_multi_pow)
    monkeypatch.setattr(pairing, "multi_pairing_batch", multi_pairing_batch)
    return calls


@pytest.mark.parametrize("nv", [3, 4])
def test_verifier_mirror_follows_the_oracle_verifier(stand_ins, nv):
    m_row = nv - nv // 2
    t = o.rand_scalars(m_row, 5100 + nv)
    ck = osp.setup_ck(t)
    vk = ver.setup_vk(t)
    z = o.rand_scalars(1 << nv, 5110 + nv)
    r = o.rand_scalars(nv, 5120 + nv)
    poly = osp.Polynomial(z)
    v = poly.eval(r)
    comm_list, T = poly.commit(ck)
    U, pst_proof, mp = poly.open(word_transcript(), comm_list, ck, r, T)
    assert ver.sqrt_pst_verify(vk, word_transcript(), U, r, v, pst_proof, mp, T) is True

    vk_e = multilinear_pc.VerifierKey(nv=vk["nv"], g=h.pts_to_np([vk["g"]])[0], h=g2_np([vk["h"]])[0],
                                      g_mask_random=h.pts_to_np(vk["g_mask"]), h_mask_random=g2_np(vk["h_mask"]))

    def mirror_proof(d):
        return mipp.MippProofG1(comms_u=[(h.pts_to_np([l])[0], h.pts_to_np([rr])[0]) for l, rr in d["comms_u"]],
                                comms_t=[(gt_np(l), gt_np(rr)) for l, rr in d["comms_t"]],
                                final_a=h.pts_to_np([d["final_a"]])[0], final_h=g2_np([d["final_h"]])[0],
                                pst_proof_h=h.pts_to_np(d["pst_proof_h"]))

    def mirror_verdict(u, point, value, pst, d, t_gt):
        return sqrt_pst.Polynomial.verify(word_transcript(), vk_e, h.pts_to_np([u])[0], point, value, g2_np(pst),
                                          mirror_proof(d), gt_np(t_gt))

    before = dict(stand_ins)
    assert mirror_verdict(U, r, v, pst_proof, mp, T) is True
    # ONE ragged G1 batch, ONE GT fold, ONE pass of pairing products per verification
    assert stand_ins["rows"] == before["rows"] + 1 and stand_ins["gt"] == before["gt"] + 1
    assert stand_ins["pairing_batches"] == before["pairing_batches"] + 1
    cases = [("value", U, r, (v + 1) % o.R_ORDER, pst_proof, mp, T)]
    bad = dict(mp)
    bad["final_a"] = o.add(bad["final_a"], o.G)
    cases.append(("final_a", U, r, v, pst_proof, bad, T))
    bad = dict(mp)
    bad["pst_proof_h"] = list(reversed(bad["pst_proof_h"])) if len(bad["pst_proof_h"]) > 1 else [o.G]
    cases.append(("pst_proof_h", U, r, v, pst_proof, bad, T))
    cases.append(("pst_proof", U, r, v, [o2.add(pst_proof[0], o2.G2)] + list(pst_proof[1:]), mp, T))
    cases.append(("T", U, r, v, pst_proof, mp, pr.f12_sqr(T)))
    cases.append(("point", U, [(r[0] + 1) % o.R_ORDER] + list(r[1:]), v, pst_proof, mp, T))
    for what, u, point, value, pst, d, t_gt in cases:
        want = ver.sqrt_pst_verify(vk, word_transcript(), u, point, value, pst, d, t_gt)
        assert want is False, what
        assert mirror_verdict(u, point, value, pst, d, t_gt) is False, what
    # the standalone halves agree with the oracle's too (honest and wrong value)
    a_rev = list(r[: nv // 2 + nv % 2])[::-1]
    for value in (v, (v + 1) % o.R_ORDER):
        assert multilinear_pc.check(vk_e, h.pts_to_np([U])[0], a_rev, value, g2_np(pst_proof)) == ver.check(vk, U, a_rev, value, pst_proof)
    b = list(r[nv // 2 + nv % 2:])
    assert mirror_proof(mp).verify(vk_e, word_transcript(), b, h.pts_to_np([U])[0], gt_np(T)) is True
