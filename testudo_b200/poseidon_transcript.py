"""Mirror of `PoseidonTranscript<F>` (src/poseidon_transcript.rs:12-125) and of the Poseidon parameter set the reference
ships (src/parameters.rs:17-185, "bls12377_rate2_constraints_fr": rate 2, capacity 1, alpha 17, 8 full + 31 partial
rounds) -- the Fiat-Shamir transcript of `Polynomial::open` / `MippProof::prove`.

* The round constants and the MDS matrix are REGENERATED, not copied: `find_poseidon_ark_and_mds` of ark-crypto-primitives
  0.4 (`sponge/poseidon/grain_lfsr.rs`, the Grain LFSR of the Poseidon paper in self-shrinking mode) restated below
  reproduces the 117 + 9 constants of src/parameters.rs exactly (tests/test_poseidon_transcript.py checks their SHA-256,
  and the file itself wherever /root/reference exists).
* The sponge is host code in the reference and here: tb200_poseidon_* in the C ABI (csrc/poseidon_host.cpp).
* `PoseidonTranscript(field="fq")` is what the commitment path uses (`PoseidonTranscript<E::BaseField>`, src/mipp.rs:32,
  src/sqrt_pst.rs:170,325 with `get_bls12377_fq_params()`: the SAME integers read as Fq elements, src/parameters.rs:309);
  `field="fr"` is `Fr::poseidon_params()` (benches/pst.rs:23).
* `as_challenge()` adapts a transcript to the `challenge(label, values)` callback of the prover mirrors
  (testudo_b200/sqrt_pst.py, mipp.py): values are appended as the reference does -- `append` = uncompressed
  `CanonicalSerialize` bytes absorbed as a `Vec<u8>` (:21-27) --, the challenge is `challenge_scalar::<Fr>` (:29-31).
"""
from __future__ import annotations

import ctypes
from functools import lru_cache
from typing import Callable, List, Sequence, Tuple

import numpy as np

from . import _lib, curve, serialize

FULL_ROUNDS, PARTIAL_ROUNDS, ALPHA, RATE, CAPACITY = 8, 31, 17, 2, 1     # src/parameters.rs:148-151,182-184


class GrainLFSR:
    """ark-crypto-primitives 0.4 `PoseidonGrainLFSR`: 80-bit state b0..b79; b0 b1 = field type (1 = prime field, stored
    as bits 0, 1), b2..b5 = S-box (x^alpha: 0; inverse: 1, stored in b5), b6..b17 = bits of the modulus size, b18..b29 =
    state width, b30..b39 = full rounds, b40..b49 = partial rounds (all MSB first), b50..b79 = 1; 160 warm-up updates;
    feedback b62 ^ b51 ^ b38 ^ b23 ^ b13 ^ b0. Output bits are taken in pairs: the second bit of a pair is emitted iff
    the first is 1."""

    def __init__(self, inverse_sbox: bool, prime_bits: int, width: int, full_rounds: int, partial_rounds: int):
        st = [False] * 80
        st[1] = True
        st[5] = bool(inverse_sbox)

        def put(value: int, lo: int, n: int) -> None:
            for i in range(n):
                st[lo + n - 1 - i] = bool((value >> i) & 1)

        put(prime_bits, 6, 12)
        put(width, 18, 12)
        put(full_rounds, 30, 10)
        put(partial_rounds, 40, 10)
        for i in range(50, 80):
            st[i] = True
        self.state, self.head, self.prime_bits = st, 0, prime_bits
        for _ in range(160):
            self._update()

    def _update(self) -> bool:
        st, h = self.state, self.head
        bit = st[(h + 62) % 80] ^ st[(h + 51) % 80] ^ st[(h + 38) % 80] ^ st[(h + 23) % 80] ^ st[(h + 13) % 80] ^ st[h]
        st[h] = bit
        self.head = (h + 1) % 80
        return bit

    def _value(self) -> int:
        v = 0
        for _ in range(self.prime_bits):            # most significant bit first
            first = self._update()
            while not first:
                self._update()
                first = self._update()
            v = (v << 1) | int(self._update())
        return v

    def rejection_sample(self, modulus: int) -> int:
        while True:
            v = self._value()
            if v < modulus:
                return v

    def mod_p(self, modulus: int) -> int:
        return self._value() % modulus


@lru_cache(maxsize=None)
def find_poseidon_ark_and_mds(prime_bits: int, modulus: int, rate: int, full_rounds: int, partial_rounds: int,
                              skip_matrices: int = 0) -> Tuple[Tuple[Tuple[int, ...], ...], Tuple[Tuple[int, ...], ...]]:
    """ark-crypto-primitives 0.4 `find_poseidon_ark_and_mds`: (ark[rounds][rate + 1], mds[rate + 1][rate + 1])."""
    width = rate + 1
    lfsr = GrainLFSR(False, prime_bits, width, full_rounds, partial_rounds)
    ark = tuple(tuple(lfsr.rejection_sample(modulus) for _ in range(width)) for _ in range(full_rounds + partial_rounds))
    for _ in range(skip_matrices):
        for _ in range(2 * width):
            lfsr.mod_p(modulus)
    xs = [lfsr.mod_p(modulus) for _ in range(width)]
    ys = [lfsr.mod_p(modulus) for _ in range(width)]
    mds = tuple(tuple(pow(xs[i] + ys[j], -1, modulus) for j in range(width)) for i in range(width))
    return ark, mds


def reference_parameters():
    """The parameter set of src/parameters.rs:17-151, generated over Fr (253 bits): (ark, mds) as integers. The Fq
    variants of the reference reuse these integers (src/parameters.rs:231-338)."""
    return find_poseidon_ark_and_mds(253, curve.R_ORDER, RATE, FULL_ROUNDS, PARTIAL_ROUNDS)


class PoseidonTranscript:
    """src/poseidon_transcript.rs:12-125 over `field` in {"fr", "fq"}."""

    def __init__(self, field: str = "fq"):
        assert field in ("fr", "fq")
        self.field = field
        self._limbs = 4 if field == "fr" else 6
        ark, mds = reference_parameters()
        lib = _lib.load()                                               # no CUDA call: works without a GPU
        a = np.array([curve.limbs64(v, self._limbs) for row in ark for v in row], dtype=np.uint64)
        m = np.array([curve.limbs64(v, self._limbs) for row in mds for v in row], dtype=np.uint64)
        self._h = ctypes.c_void_p()
        _lib.check(lib.tb200_poseidon_new(0 if field == "fr" else 1, FULL_ROUNDS, PARTIAL_ROUNDS, ALPHA, RATE, CAPACITY,
                                          a.ctypes.data_as(ctypes.c_void_p), m.ctypes.data_as(ctypes.c_void_p),
                                          ctypes.byref(self._h)))
        self._lib = lib

    # -- Transcript trait (src/transcript.rs:7-14) ---------------------------------------------------------------------
    def domain_sep(self) -> None:                                       # :17-19
        self.append_bytes(b"", b"testudo")

    def append(self, _label: bytes, value) -> None:
        """:21-27 -- `value` is a C-ABI word array: [12] G1, [24] G2, [72] GT, [4] Fr (Montgomery limbs)."""
        w = np.ascontiguousarray(value, dtype=np.uint64).reshape(-1)
        if len(w) not in (4, 12, 24, 72):
            raise ValueError(f"cannot append a value of {len(w)} words")
        # the encoding (`encode_uncompressed` below) and the absorb in one library call: the Python big integers cost
        # ~0.1 ms per MIPP round
        _lib.check(self._lib.tb200_poseidon_append_words(self._h, w.ctypes.data_as(ctypes.c_void_p), len(w)))

    def challenge_scalar(self, _label: bytes = b"") -> int:            # :29-31
        out = np.zeros(4, dtype=np.uint64)
        _lib.check(self._lib.tb200_poseidon_squeeze_fr(self._h, out.ctypes.data_as(ctypes.c_void_p)))
        return curve.from_limbs64(out) % curve.R_ORDER

    def challenge_scalar_vec(self, label: bytes, n: int) -> List[int]:  # src/transcript.rs:11-13
        return [self.challenge_scalar(label) for _ in range(n)]

    # -- inherent methods (:63-121) -------------------------------------------------------------------------------------------
    def append_bytes(self, _label: bytes, data: bytes) -> None:        # :68-70
        buf = (ctypes.c_uint8 * max(len(data), 1)).from_buffer_copy(data.ljust(1, b"\0"))
        _lib.check(self._lib.tb200_poseidon_absorb_bytes(self._h, buf, len(data)))

    def append_u64(self, _label: bytes, x: int) -> None:               # :64-66: absorb(&u64) = one native element
        self._absorb_native([x])

    def append_scalar(self, _label: bytes, scalar: int) -> None:       # :72-74 (native field only, as `Absorb for F`)
        self._absorb_native([scalar])

    def append_scalar_vector(self, label: bytes, scalars: Sequence[int]) -> None:   # :89-97
        for s in scalars:
            self.append_scalar(label, s)

    def append_point(self, _label: bytes, words) -> None:              # :76-87, 115-121: COMPRESSED encoding
        w = np.asarray(words, dtype=np.uint64).reshape(-1)
        self.append_bytes(b"", serialize.g1_bytes(w, True) if len(w) == 12 else serialize.g2_bytes(w, True))

    append_g1 = append_point                                            # :107-114

    def append_gt(self, _label: bytes, words) -> None:                 # :99-105 (Fq12 has no compressed form)
        self.append_bytes(b"", serialize.gt_bytes(words))

    def new_from_state(self, challenge: int) -> None:                  # :51-54
        _lib.check(self._lib.tb200_poseidon_reset(self._h))
        self.append_scalar(b"", challenge)

    def _absorb_native(self, values: Sequence[int]) -> None:
        a = np.array([curve.limbs64(int(v), self._limbs) for v in values], dtype=np.uint64)
        _lib.check(self._lib.tb200_poseidon_absorb_native(self._h, a.ctypes.data_as(ctypes.c_void_p), len(values)))

    def squeeze_native(self, n: int) -> List[int]:
        out = np.zeros((n, self._limbs), dtype=np.uint64)
        _lib.check(self._lib.tb200_poseidon_squeeze_native(self._h, out.ctypes.data_as(ctypes.c_void_p), n))
        return [curve.from_limbs64(r) for r in out]

    def as_challenge(self) -> Callable[[bytes, list], int]:
        """The `challenge(label, values)` callback of `Polynomial.open` / `MippProofG1.prove`: append every value the
        reference appends at that point, then squeeze -- except for the label b"U", which is an append without a
        challenge (src/mipp.rs:56)."""
        def challenge(label: bytes, values) -> int:
            for v in values:
                self.append(label, v)
            if label == b"U":
                return 0
            return self.challenge_scalar(label)
        return challenge

    def close(self) -> None:
        if self._h is not None:
            self._lib.tb200_poseidon_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def encode_uncompressed(value) -> bytes:
    """`serialize_with_mode(.., Compress::No)` of a C-ABI word array, by its length."""
    w = np.asarray(value, dtype=np.uint64).reshape(-1)
    if len(w) == 12:
        return serialize.g1_bytes(w, False)
    if len(w) == 24:
        return serialize.g2_bytes(w, False)
    if len(w) == 72:
        return serialize.gt_bytes(w)
    if len(w) == 4:
        return serialize.fr_bytes(w)
    raise ValueError(f"cannot append a value of {len(w)} words")
