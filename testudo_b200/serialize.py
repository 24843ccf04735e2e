"""Wire formats of the values the commitment path emits: ark-serialize 0.4 `CanonicalSerialize`
for `G1Affine`, `G2Affine`, `Fq12` (GT), `Fr` and the proof structs built from them -- what the reference writes with
`serialize_with_mode(.., Compress::Yes)` to measure proof sizes (benches/pst.rs:64-74) and with `Compress::No` when a
value is appended to the transcript (src/poseidon_transcript.rs:22-28).

Host-side byte shuffling only (the reference does this on the CPU too); inputs are the C ABI's numpy word arrays
(Montgomery limbs), outputs are `bytes`.

Format restated from ark-serialize / ark-ec 0.4 (un-vendored dependency; cannot be checked against a binary here, see
DESIGN.md "parity unpinned"):
  * Fp element: canonical (non-Montgomery) value, little-endian, ceil((bits + flag bits) / 8) bytes: Fq -> 48, Fr -> 32.
  * Fq2: c0 then c1;  Fq12: the twelve Fq coefficients in tower order (576 bytes, no flags).
  * SW affine point, compressed: x with `SWFlags` OR-ed into the top bits of the LAST byte:
        bit 7 = y is "negative" (y > -y; Fq compares canonical integers, Fq2 compares c1 first, then c0),
        bit 6 = point at infinity (x written as zero).
    uncompressed: x, then y carrying the same flags. G1: 48 / 96 bytes, G2: 96 / 192 bytes.
  * Vec<T>: u64 little-endian length, then the elements; usize fields as u64; tuples / structs field by field.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import numpy as np

from . import curve

Q = curve.Q
_FQ_RINV = pow(curve.FQ_R, -1, Q)
_FR_RINV = pow(curve.FR_R, -1, curve.R_ORDER)
FLAG_NEG = 1 << 7
FLAG_INF = 1 << 6
TWIST_B = (0, (-pow(5, -1, Q)) % Q)   # B' = 1/u of the G2 curve y^2 = x^3 + B'


def _fq_vals(words, count: int) -> List[int]:
    w = np.asarray(words, dtype=np.uint64).reshape(count, 6)
    return [curve.from_limbs64(row) * _FQ_RINV % Q for row in w]


def _fq_words(vals: Sequence[int]) -> np.ndarray:
    return np.array([curve.limbs64(v % Q * curve.FQ_R % Q, 6) for v in vals], dtype=np.uint64).reshape(-1)


def fr_bytes(words) -> bytes:
    v = curve.from_limbs64(np.asarray(words, dtype=np.uint64).reshape(4)) * _FR_RINV % curve.R_ORDER
    return v.to_bytes(32, "little")


def _neg_flag(y: Sequence[int]) -> int:
    """`y <= -y` -> positive. Lexicographic from the highest coefficient (QuadExtField's Ord)."""
    ny = [(-c) % Q for c in y]
    return FLAG_NEG if tuple(reversed(y)) > tuple(reversed(ny)) else 0


def _point_bytes(words, ncoord: int, compress: bool) -> bytes:
    w = np.asarray(words, dtype=np.uint64).reshape(-1)
    assert len(w) == 12 * ncoord
    if not w.any():                                        # identity (the C ABI's all-zero encoding)
        out = bytearray(48 * ncoord * (1 if compress else 2))
        out[-1] |= FLAG_INF
        return bytes(out)
    c = _fq_vals(w, 2 * ncoord)
    x, y = c[:ncoord], c[ncoord:]
    body = b"".join(v.to_bytes(48, "little") for v in (x if compress else x + y))
    out = bytearray(body)
    out[-1] |= _neg_flag(y)
    return bytes(out)


def g1_bytes(words, compress: bool = True) -> bytes:
    return _point_bytes(words, 1, compress)


def g2_bytes(words, compress: bool = True) -> bytes:
    return _point_bytes(words, 2, compress)


def gt_bytes(words) -> bytes:
    return b"".join(v.to_bytes(48, "little") for v in _fq_vals(words, 12))


def vec_bytes(items: Sequence[bytes]) -> bytes:
    return len(items).to_bytes(8, "little") + b"".join(items)


# ---- deserialisation (compressed points need a square root: Fq has q = 1 mod 4 -> Tonelli-Shanks) ------------------------
def _fq_sqrt(a: int) -> Optional[int]:
    a %= Q
    if a == 0:
        return 0
    if pow(a, (Q - 1) // 2, Q) != 1:
        return None
    s, t = 0, Q - 1
    while t % 2 == 0:
        s, t = s + 1, t // 2
    z = 2
    while pow(z, (Q - 1) // 2, Q) != Q - 1:
        z += 1
    m, c, tt, r = s, pow(z, t, Q), pow(a, t, Q), pow(a, (t + 1) // 2, Q)
    while tt != 1:
        i, x = 0, tt
        while x != 1:
            x = x * x % Q
            i += 1
        b = pow(c, 1 << (m - i - 1), Q)
        m, c = i, b * b % Q
        tt, r = tt * c % Q, r * b % Q
    return r


def _fq2_mul(a, b):
    return ((a[0] * b[0] - 5 * a[1] * b[1]) % Q, (a[0] * b[1] + a[1] * b[0]) % Q)


def _fq2_sqrt(a) -> Optional[Tuple[int, int]]:
    """sqrt in Fq[u]/(u^2 + 5) via the norm: a = (x + y u)^2 with x^2 = (a0 +- sqrt(N(a))) / 2."""
    a0, a1 = a[0] % Q, a[1] % Q
    if a1 == 0:
        r = _fq_sqrt(a0)
        if r is not None:
            return (r, 0)
        r = _fq_sqrt(a0 * pow(-5, -1, Q) % Q)          # a0 = -5 y^2
        return None if r is None else (0, r)
    n = _fq_sqrt((a0 * a0 + 5 * a1 * a1) % Q)
    if n is None:
        return None
    inv2 = pow(2, -1, Q)
    for cand in ((a0 + n) * inv2 % Q, (a0 - n) * inv2 % Q):
        x = _fq_sqrt(cand)
        if x is not None and x != 0:
            y = a1 * pow(2 * x, -1, Q) % Q
            if _fq2_mul((x, y), (x, y)) == (a0, a1):
                return (x, y)
    return None


def _point_from_bytes(data: bytes, ncoord: int, compress: bool) -> np.ndarray:
    size = 48 * ncoord * (1 if compress else 2)
    if len(data) != size:
        raise ValueError(f"expected {size} bytes, got {len(data)}")
    flags = data[-1] & (FLAG_NEG | FLAG_INF)
    if flags == (FLAG_NEG | FLAG_INF):
        raise ValueError("invalid flag combination")
    buf = bytearray(data)
    buf[-1] &= 0x3F
    vals = [int.from_bytes(buf[48 * i: 48 * i + 48], "little") for i in range(len(buf) // 48)]
    if any(v >= Q for v in vals):
        raise ValueError("coordinate not reduced")
    if flags & FLAG_INF:
        return np.zeros(12 * ncoord, dtype=np.uint64)
    x = vals[:ncoord]
    if compress:
        if ncoord == 1:
            r = _fq_sqrt((pow(x[0], 3, Q) + 1) % Q)
            y = None if r is None else [r]
        else:
            x3 = _fq2_mul(_fq2_mul(x, x), x)
            r = _fq2_sqrt(((x3[0] + TWIST_B[0]) % Q, (x3[1] + TWIST_B[1]) % Q))
            y = None if r is None else list(r)
        if y is None:
            raise ValueError("x is not on the curve")
        if bool(_neg_flag(y)) != bool(flags & FLAG_NEG):
            y = [(-c) % Q for c in y]
    else:
        y = vals[ncoord:]
    return _fq_words(list(x) + list(y))


def g1_from_bytes(data: bytes, compress: bool = True) -> np.ndarray:
    return _point_from_bytes(data, 1, compress)


def g2_from_bytes(data: bytes, compress: bool = True) -> np.ndarray:
    return _point_from_bytes(data, 2, compress)


def gt_from_bytes(data: bytes) -> np.ndarray:
    if len(data) != 576:
        raise ValueError("expected 576 bytes")
    return _fq_words([int.from_bytes(data[48 * i: 48 * i + 48], "little") for i in range(12)])


# ---- the reference's proof structs ----------------------------------------------------------------------------------
def commitment_bytes(nv: int, g_product, compress: bool = True) -> bytes:
    """`Commitment<E>{nv: usize, g_product: G1Affine}` (ark-poly-commit)."""
    return int(nv).to_bytes(8, "little") + g1_bytes(g_product, compress)


def pst_proof_bytes(proofs, compress: bool = True) -> bytes:
    """`Proof<E>{proofs: Vec<G2Affine>}`."""
    return vec_bytes([g2_bytes(p, compress) for p in np.asarray(proofs, dtype=np.uint64).reshape(-1, 24)])


def mipp_proof_bytes(comms_t, comms_u, final_a, final_h, pst_proof_h, compress: bool = True) -> bytes:
    """`MippProof<E>` field by field (src/mipp.rs:21-28): comms_t, comms_u, final_a, final_h, pst_proof_h."""
    out = vec_bytes([gt_bytes(l) + gt_bytes(r) for l, r in comms_t])
    out += vec_bytes([g1_bytes(l, compress) + g1_bytes(r, compress) for l, r in comms_u])
    out += g1_bytes(final_a, compress) + g2_bytes(final_h, compress)
    out += vec_bytes([g1_bytes(p, compress) for p in np.asarray(pst_proof_h, dtype=np.uint64).reshape(-1, 12)])
    return out


# ---- `CanonicalDeserialize` of the same structs: what a verifier receives ------------------------------------------------
def _take(data: bytes, pos: int, n: int) -> Tuple[bytes, int]:
    if pos + n > len(data):
        raise ValueError("truncated input")
    return data[pos:pos + n], pos + n


def _take_len(data: bytes, pos: int) -> Tuple[int, int]:
    raw, pos = _take(data, pos, 8)
    return int.from_bytes(raw, "little"), pos


def commitment_from_bytes(data: bytes, compress: bool = True) -> Tuple[int, np.ndarray]:
    """-> (nv, g_product)"""
    nv, pos = _take_len(data, 0)
    raw, pos = _take(data, pos, 48 if compress else 96)
    if pos != len(data):
        raise ValueError("trailing bytes")
    return nv, g1_from_bytes(raw, compress)


def pst_proof_from_bytes(data: bytes, compress: bool = True) -> np.ndarray:
    """`Proof<E>` -> [nv, 24]"""
    n, pos = _take_len(data, 0)
    size = 96 if compress else 192
    out = np.zeros((n, 24), dtype=np.uint64)
    for i in range(n):
        raw, pos = _take(data, pos, size)
        out[i] = g2_from_bytes(raw, compress)
    if pos != len(data):
        raise ValueError("trailing bytes")
    return out


def mipp_proof_from_bytes(data: bytes, compress: bool = True):
    """`MippProof<E>` (src/mipp.rs:21-28) -> (comms_t, comms_u, final_a, final_h, pst_proof_h) as `mipp_proof_bytes` takes
    them: lists of (left, right) pairs, [12], [24], [m, 12]."""
    g1 = 48 if compress else 96
    n, pos = _take_len(data, 0)
    comms_t = []
    for _ in range(n):
        l, pos = _take(data, pos, 576)
        r, pos = _take(data, pos, 576)
        comms_t.append((gt_from_bytes(l), gt_from_bytes(r)))
    n, pos = _take_len(data, pos)
    comms_u = []
    for _ in range(n):
        l, pos = _take(data, pos, g1)
        r, pos = _take(data, pos, g1)
        comms_u.append((g1_from_bytes(l, compress), g1_from_bytes(r, compress)))
    raw, pos = _take(data, pos, g1)
    final_a = g1_from_bytes(raw, compress)
    raw, pos = _take(data, pos, 2 * g1)
    final_h = g2_from_bytes(raw, compress)
    n, pos = _take_len(data, pos)
    pst_proof_h = np.zeros((n, 12), dtype=np.uint64)
    for i in range(n):
        raw, pos = _take(data, pos, g1)
        pst_proof_h[i] = g1_from_bytes(raw, compress)
    if pos != len(data):
        raise ValueError("trailing bytes")
    return comms_t, comms_u, final_a, final_h, pst_proof_h


def proof_size(pst_proof, mipp) -> int:
    """`br.proof_size = p1.len() + p2.len()` of benches/pst.rs:64-74 (both compressed); `mipp` is a MippProofG1."""
    p1 = pst_proof_bytes(pst_proof, True)
    p2 = mipp_proof_bytes(mipp.comms_t, mipp.comms_u, mipp.final_a, mipp.final_h, mipp.pst_proof_h, True)
    return len(p1) + len(p2)
