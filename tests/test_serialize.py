"""CPU tests for the wire formats (testudo_b200/serialize.py): sizes, flag bits, the `y <= -y`
sign rule, infinity, and decompression round trips against the oracle's curve arithmetic."""
import random

import numpy as np
import pytest

import helpers as h
from oracle import bls12_377 as o
from oracle import bls12_377_g2 as o2
from oracle import pairing as pr
from testudo_b200 import serialize as ser


def g2w(p):
    return np.array(o2.affine_to_words(p), dtype=np.uint64)


def test_g1_compressed_and_uncompressed():
    pts, _ = o.rand_points(12, 71)
    for p in pts + [o.G, o.neg(o.G)]:
        w = h.pts_to_np([p])[0]
        c, u = ser.g1_bytes(w, True), ser.g1_bytes(w, False)
        assert len(c) == 48 and len(u) == 96
        x = int.from_bytes(c, "little") & ((1 << 382) - 1)
        assert x == p[0] and u[:48] == p[0].to_bytes(48, "little")
        neg = p[1] > (o.Q - p[1]) % o.Q
        assert bool(c[-1] & 0x80) == neg and bool(u[-1] & 0x80) == neg and not (c[-1] & 0x40)
        assert int.from_bytes(u[48:], "little") & ((1 << 382) - 1) == p[1]
        assert np.array_equal(ser.g1_from_bytes(c, True), w)
        assert np.array_equal(ser.g1_from_bytes(u, False), w)
    # a point and its negation differ only in the sign flag
    a, b = ser.g1_bytes(h.pts_to_np([o.G])[0]), ser.g1_bytes(h.pts_to_np([o.neg(o.G)])[0])
    assert a[:-1] == b[:-1] and (a[-1] ^ b[-1]) == 0x80
    inf = ser.g1_bytes(np.zeros(12, np.uint64))
    assert inf == bytes(47) + b"\x40" and not ser.g1_from_bytes(inf).any()
    with pytest.raises(ValueError):
        ser.g1_from_bytes(bytes(47) + b"\xc0")


def test_g2_compressed_and_uncompressed():
    pts, _ = o2.rand_points(8, 72)
    for p in pts + [o2.G2, o2.neg(o2.G2)]:
        w = g2w(p)
        c, u = ser.g2_bytes(w, True), ser.g2_bytes(w, False)
        assert len(c) == 96 and len(u) == 192
        assert c[:48] == p[0][0].to_bytes(48, "little")
        y, ny = p[1], o2.f2_neg(p[1])
        neg = (y[1], y[0]) > (ny[1], ny[0])                     # QuadExtField orders by c1, then c0
        assert bool(c[-1] & 0x80) == neg and not (c[-1] & 0x40)
        assert np.array_equal(ser.g2_from_bytes(c, True), w)
        assert np.array_equal(ser.g2_from_bytes(u, False), w)
    inf = ser.g2_bytes(np.zeros(24, np.uint64))
    assert inf == bytes(95) + b"\x40" and not ser.g2_from_bytes(inf).any()


def test_gt_fr_and_proof_structs():
    rng = random.Random(73)
    f = tuple((rng.randrange(o.Q), rng.randrange(o.Q)) for _ in range(6))
    w = np.array(pr.to_words(f), dtype=np.uint64)
    b = ser.gt_bytes(w)
    assert len(b) == 576
    # tower order c0.c0, c0.c1, c0.c2, c1.c0, c1.c1, c1.c2 == flat coefficients 0, 2, 4, 1, 3, 5
    assert b[:48] == f[0][0].to_bytes(48, "little") and b[96:144] == f[2][0].to_bytes(48, "little")
    assert np.array_equal(ser.gt_from_bytes(b), w)
    s = 0x1234567890ABCDEF
    assert ser.fr_bytes(h.scalars_to_np([s], mont=True)[0]) == s.to_bytes(32, "little")
    pts, _ = o.rand_points(4, 74)
    qs, _ = o2.rand_points(3, 75)
    g1w = h.pts_to_np(pts)
    pst = np.array([o2.affine_to_words(q) for q in qs], dtype=np.uint64)
    assert len(ser.pst_proof_bytes(pst)) == 8 + 3 * 96
    assert len(ser.commitment_bytes(5, g1w[0])) == 8 + 48
    mp = ser.mipp_proof_bytes([(w, w)], [(g1w[0], g1w[1])], g1w[2], pst[0], g1w[:1])
    assert len(mp) == (8 + 2 * 576) + (8 + 2 * 48) + 48 + 96 + (8 + 48)
    assert mp[:8] == (1).to_bytes(8, "little")


@pytest.mark.parametrize("compress", [True, False])
def test_proof_structs_round_trip_through_bytes(compress):
    """`CanonicalDeserialize` of `Commitment`, `Proof` and `MippProof` (what `Polynomial::verify` receives, benches/pst.rs:
    64-90): bytes -> words -> bytes is the identity, identity points included; truncated and over-long inputs are errors."""
    from testudo_b200.mipp import MippProofG1

    rng = random.Random(77)
    gts = [np.array(pr.to_words(tuple((rng.randrange(o.Q), rng.randrange(o.Q)) for _ in range(6))), dtype=np.uint64)
           for _ in range(4)]
    pts, _ = o.rand_points(7, 78)
    pts[3] = None                                           # an identity among comms_u
    qs, _ = o2.rand_points(4, 79)
    g1w = h.pts_to_np(pts)
    g2s = np.array([o2.affine_to_words(q) for q in qs], dtype=np.uint64)
    proof = MippProofG1(comms_u=[(g1w[0], g1w[1]), (g1w[2], g1w[3])], comms_t=[(gts[0], gts[1]), (gts[2], gts[3])],
                        final_a=g1w[4], final_h=g2s[0], pst_proof_h=g1w[5:7])
    data = proof.to_bytes(compress)
    back = MippProofG1.from_bytes(data, compress)
    assert back.to_bytes(compress) == data
    assert all(np.array_equal(a, b) for pa, pb in zip(back.comms_u, proof.comms_u) for a, b in zip(pa, pb))
    assert all(np.array_equal(a, b) for pa, pb in zip(back.comms_t, proof.comms_t) for a, b in zip(pa, pb))
    assert np.array_equal(back.final_a, g1w[4]) and np.array_equal(back.final_h, g2s[0])
    assert np.array_equal(back.pst_proof_h, g1w[5:7])
    pst = ser.pst_proof_bytes(g2s[1:], compress)
    assert np.array_equal(ser.pst_proof_from_bytes(pst, compress), g2s[1:])
    nv, gp = ser.commitment_from_bytes(ser.commitment_bytes(13, g1w[6], compress), compress)
    assert nv == 13 and np.array_equal(gp, g1w[6])
    for bad in (data[:-1], data + b"\0"):
        with pytest.raises(ValueError):
            MippProofG1.from_bytes(bad, compress)
    with pytest.raises(ValueError):
        ser.pst_proof_from_bytes(pst[:-3], compress)
