"""Poseidon transcript (src/poseidon_transcript.rs:12-125, src/parameters.rs:17-185): the regenerated
parameters against the reference's constants, the C++ sponge of the C ABI against the independent Python restatement,
and the transcript mirror's encodings. CPU only: the sponge is host code (no GPU needed, no CUDA call made)."""
import hashlib
import os
import re

import numpy as np
import pytest

import helpers as h
from oracle import bls12_377 as o
from oracle import poseidon as op
from testudo_b200 import curve, poseidon_transcript as pt

# SHA-256 over the decimal strings of the 117 round constants followed by the 9 MDS entries, "\n"-joined, computed from
# /root/reference/src/parameters.rs:20-146 in the build container (tests/golden/make_poseidon_digest.py)
REFERENCE_CONSTANTS_SHA256 = open(os.path.join(h.GOLDEN_DIR, "poseidon_constants.sha256")).read().split()[0]


def _digest(ark, mds) -> str:
    vals = [str(v) for row in ark for v in row] + [str(v) for row in mds for v in row]
    return hashlib.sha256("\n".join(vals).encode()).hexdigest()


def test_regenerated_parameters_reproduce_the_reference_constants():
    ark, mds = pt.reference_parameters()
    assert len(ark) == 39 and all(len(r) == 3 for r in ark) and len(mds) == 3
    assert _digest(ark, mds) == REFERENCE_CONSTANTS_SHA256
    ref = "/root/reference/src/parameters.rs"
    if os.path.exists(ref):                                    # the build container: compare with the file itself
        src = open(ref).read()
        blk = src[src.index("pub static ref FR"):src.index('"rate" => 2')]
        nums = [int(x) for x in re.findall(r'"(\d{20,})"', blk)]
        assert nums == [v for row in ark for v in row] + [v for row in mds for v in row]


def _oracle_sponge(field):
    ark, mds = pt.reference_parameters()
    p = o.R_ORDER if field == "fr" else curve.Q
    return op.PoseidonSponge(p, pt.FULL_ROUNDS, pt.PARTIAL_ROUNDS, pt.ALPHA, mds, ark, pt.RATE, pt.CAPACITY)


@pytest.mark.parametrize("field", ["fr", "fq"])
def test_cpp_sponge_equals_python_restatement(field):
    rng = np.random.default_rng(11 if field == "fr" else 12)
    t = pt.PoseidonTranscript(field)
    s = _oracle_sponge(field)
    p = o.R_ORDER if field == "fr" else curve.Q
    for step in range(60):
        kind = int(rng.integers(0, 4))
        if kind == 0:
            data = rng.bytes(int(rng.integers(0, 700)))
            t.append_bytes(b"", data)
            s.absorb_bytes(data)
        elif kind == 1:
            v = int.from_bytes(rng.bytes(64), "little") % p
            t.append_scalar(b"", v)
            s.absorb_elements([v])
        elif kind == 2:
            n = int(rng.integers(1, 6))
            assert t.squeeze_native(n) == s.squeeze_native(n), step
        else:
            want = s.squeeze_native(1)[0] if field == "fr" else s.squeeze_foreign(o.R_ORDER)
            assert t.challenge_scalar(b"c") == want, step


@pytest.mark.parametrize("field", ["fr", "fq"])
def test_cpp_sponge_on_extreme_field_elements(field):
    """The host multiplier (interleaved CIOS on two carry words, three-term dot product with one reduction) on operands at
    the edges: p - 1, p - 2, 2^k - 1, all-ones limbs -- long absorb / squeeze chains against the Python restatement."""
    import random
    rng = random.Random(5 if field == "fr" else 6)
    t = pt.PoseidonTranscript(field)
    s = _oracle_sponge(field)
    p = o.R_ORDER if field == "fr" else curve.Q
    edge = [p - 1, p - 2, 1, 0, (1 << (p.bit_length() - 1)) - 1, (1 << 64) - 1, ((1 << 128) - 1) << 64, p >> 1, (p >> 1) + 1]
    edge += [(rng.randrange(p) | ((1 << 64) - 1)) % p for _ in range(8)]
    for rep in range(8):
        vals = edge[rep:] + edge[:rep]
        t.append_scalar_vector(b"", vals)
        s.absorb_elements(vals)
        assert t.squeeze_native(3) == s.squeeze_native(3), rep


def test_transcript_encodings_and_flow():
    """The appends of one MIPP round (src/mipp.rs:56,97-101): uncompressed G1 (96 B), GT (576 B), then a challenge."""
    from oracle import bls12_377_g2 as o2
    from oracle import pairing as pr

    u = h.pts_to_np([o.mul(1234567, o.G)])[0]
    gt = np.array(pr.to_words(pr.pairing(o.G, o2.G2)), dtype=np.uint64)
    t = pt.PoseidonTranscript("fq")
    ch = t.as_challenge()
    assert ch(b"U", [u]) == 0
    c1 = ch(b"challenge_i", [u, u, gt, gt])
    s = _oracle_sponge("fq")
    ub = pt.encode_uncompressed(u)
    gb = pt.encode_uncompressed(gt)
    assert len(ub) == 96 and len(gb) == 576
    x, y = o.mul(1234567, o.G)
    assert ub[:48] == x.to_bytes(48, "little") and ub[48:95] == y.to_bytes(48, "little")[:47]
    assert (ub[95] & 0x80) == (0x80 if y > (o.Q - y) else 0) and not (ub[95] & 0x40)
    for b in (ub, ub, ub, gb, gb):
        s.absorb_bytes(b)
    assert c1 == s.squeeze_foreign(o.R_ORDER) and 0 < c1 < o.R_ORDER
    # deterministic, and sensitive to every appended byte
    t2 = pt.PoseidonTranscript("fq")
    ch2 = t2.as_challenge()
    ch2(b"U", [u])
    assert ch2(b"challenge_i", [u, u, gt, gt]) == c1
    t3 = pt.PoseidonTranscript("fq")
    ch3 = t3.as_challenge()
    ch3(b"U", [h.pts_to_np([o.mul(1234568, o.G)])[0]])
    assert ch3(b"challenge_i", [u, u, gt, gt]) != c1
    # domain separator and the identity's encoding
    t4 = pt.PoseidonTranscript("fr")
    t4.domain_sep()
    s4 = _oracle_sponge("fr")
    s4.absorb_bytes(b"testudo")
    assert t4.challenge_scalar() == s4.squeeze_native(1)[0]
    inf = pt.encode_uncompressed(np.zeros(12, dtype=np.uint64))
    assert len(inf) == 96 and inf[95] == 0x40 and not any(inf[:95])


def test_append_words_equals_the_python_encoding():
    """tb200_poseidon_append_words (the C++ restatement of the uncompressed `CanonicalSerialize` encodings) against
    `encode_uncompressed` + absorb_bytes: Fr, G1, G2 (both signs of y, y.c1 = 0 with either sign of c0), GT, the identities
    -- the two transcripts must squeeze the same challenges."""
    import random
    from oracle import bls12_377_g2 as o2
    from oracle import pairing as pr
    from testudo_b200 import serialize
    import helpers as h

    rng = random.Random(8)
    pts, _ = o.rand_points(4, 21)
    qs, _ = o2.rand_points(4, 22)
    values = []
    for p in pts:
        values += [h.pts_to_np([p])[0], h.pts_to_np([o.neg(p)])[0]]
    for q in qs:
        values += [np.array(o2.affine_to_words(q), dtype=np.uint64), np.array(o2.affine_to_words(o2.neg(q)), dtype=np.uint64)]
    # a G2-shaped value whose y has c1 = 0 (not on the curve: only the encoder's sign rule is exercised)
    for c0 in (5, curve.Q - 5):
        values.append(serialize._fq_words([3, 4, c0, 0]))
    values += [np.zeros(12, dtype=np.uint64), np.zeros(24, dtype=np.uint64)]
    values += [np.array(pr.to_words(tuple((rng.randrange(o.Q), rng.randrange(o.Q)) for _ in range(6))), dtype=np.uint64)]
    values += [h.scalars_to_np([rng.randrange(o.R_ORDER)], mont=True)[0], h.scalars_to_np([0], mont=True)[0]]
    for field in ("fq", "fr"):
        a, b = pt.PoseidonTranscript(field), pt.PoseidonTranscript(field)
        for i, v in enumerate(values):
            a.append(b"", v)
            b.append_bytes(b"", pt.encode_uncompressed(v))
            if i % 3 == 2:
                assert a.challenge_scalar(b"") == b.challenge_scalar(b""), (field, i)
        assert a.squeeze_native(2) == b.squeeze_native(2)
    with pytest.raises(ValueError):
        pt.PoseidonTranscript("fq").append(b"", np.zeros(5, dtype=np.uint64))
