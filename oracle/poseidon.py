"""TEST INFRASTRUCTURE ONLY (never imported by testudo_b200/): the Poseidon sponge of ark-crypto-primitives 0.4
(`sponge::poseidon::PoseidonSponge`, dependency of the reference: Cargo.toml:28) with Python integers -- the checker for
the C++ sponge behind tb200_poseidon_* (csrc/poseidon_host.cpp) and for the transcript mirror
(testudo_b200/poseidon_transcript.py; reference: src/poseidon_transcript.rs:12-125).

Written independently of the C++ file from the same published source; PARITY UNPINNED against the arkworks binary
(the dependency is not vendored and there is no Rust toolchain here). The round constants it is used with
are pinned: they reproduce src/parameters.rs (tests/test_poseidon_transcript.py).
"""
from __future__ import annotations

from typing import List, Sequence


class PoseidonSponge:
    def __init__(self, modulus: int, full_rounds: int, partial_rounds: int, alpha: int, mds, ark, rate: int, capacity: int):
        self.p, self.rf, self.rp, self.alpha = modulus, full_rounds, partial_rounds, alpha
        self.mds, self.ark, self.rate, self.cap = mds, ark, rate, capacity
        self.bits = modulus.bit_length()
        self.state = [0] * (rate + capacity)
        self.mode, self.index = "absorbing", 0

    def permute(self) -> None:
        half = self.rf // 2
        st = self.state
        for r in range(self.rf + self.rp):
            st = [(s + c) % self.p for s, c in zip(st, self.ark[r])]
            if r < half or r >= half + self.rp:
                st = [pow(s, self.alpha, self.p) for s in st]
            else:
                st[0] = pow(st[0], self.alpha, self.p)
            st = [sum(st[j] * self.mds[i][j] for j in range(len(st))) % self.p for i in range(len(st))]
        self.state = st

    def _absorb_internal(self, start: int, elems: Sequence[int]) -> None:
        rem = list(elems)
        while True:
            if start + len(rem) <= self.rate:
                for i, e in enumerate(rem):
                    k = self.cap + start + i
                    self.state[k] = (self.state[k] + e) % self.p
                self.mode, self.index = "absorbing", start + len(rem)
                return
            n = self.rate - start
            for i in range(n):
                k = self.cap + start + i
                self.state[k] = (self.state[k] + rem[i]) % self.p
            self.permute()
            rem = rem[n:]
            start = 0

    def absorb_elements(self, elems: Sequence[int]) -> None:
        if not elems:
            return
        if self.mode == "absorbing":
            start = self.index
            if start == self.rate:
                self.permute()
                start = 0
            self._absorb_internal(start, elems)
        else:
            self.permute()
            self._absorb_internal(0, elems)

    def absorb_bytes(self, data: bytes) -> None:
        """`absorb(&Vec<u8>)`: `u8::batch_to_sponge_field_elements` = pack(le64(len) || bytes)."""
        buf = len(data).to_bytes(8, "little") + bytes(data)
        chunk = (self.bits - 1) // 8
        self.absorb_elements([int.from_bytes(buf[i:i + chunk], "little") for i in range(0, len(buf), chunk)])

    def squeeze_native(self, n: int) -> List[int]:
        if n == 0:
            return []
        if self.mode == "absorbing":
            self.permute()
            start = 0
        else:
            start = self.index
            if start == self.rate:
                self.permute()
                start = 0
        out: List[int] = []
        remaining = n
        while True:
            if start + remaining <= self.rate:
                out += self.state[self.cap + start: self.cap + start + remaining]
                self.mode, self.index = "squeezing", start + remaining
                return out
            k = self.rate - start
            out += self.state[self.cap + start: self.cap + start + k]
            if remaining != self.rate:          # ark tests the length BEFORE cutting off what was just read
                self.permute()
            remaining -= k
            start = 0

    def squeeze_bits(self, num_bits: int) -> List[int]:
        usable = self.bits - 1
        cnt = -(-num_bits // usable)
        bits: List[int] = []
        for e in self.squeeze_native(cnt):
            bits += [(e >> i) & 1 for i in range(usable)]
        return bits[:num_bits]

    def squeeze_foreign(self, modulus: int) -> int:
        """`squeeze_field_elements::<F2>(1)` for F2 != F: FieldElementSize::Full = MODULUS_BIT_SIZE - 1 bits,
        `from_le_bytes_mod_order` of the little-endian bit string."""
        bits = self.squeeze_bits(modulus.bit_length() - 1)
        return sum(b << i for i, b in enumerate(bits)) % modulus


class OracleTranscript:
    """`PoseidonTranscript<Fq>` on the checker's side (src/poseidon_transcript.rs:16-31): values arrive as the oracle's own
    types, tagged ("g1", affine) / ("g2", affine) / ("gt", Fq12 tuple), and are serialised here with Python integers
    (ark-serialize 0.4 uncompressed: x, then y with SWFlags in the top bits of the last byte) -- independent of
    testudo_b200/serialize.py."""

    def __init__(self, ark, mds, full_rounds=8, partial_rounds=31, alpha=17, rate=2, capacity=1):
        from . import bls12_377 as g1

        self.q, self.r = g1.Q, g1.R_ORDER
        self.sponge = PoseidonSponge(self.q, full_rounds, partial_rounds, alpha, mds, ark, rate, capacity)

    def _fq(self, v: int) -> bytes:
        return (v % self.q).to_bytes(48, "little")

    def _point(self, pt, ncoord: int) -> bytes:
        if pt is None:
            out = bytearray(96 * ncoord)
            out[-1] |= 0x40
            return bytes(out)
        x, y = pt
        xs = [x] if ncoord == 1 else list(x)
        ys = [y] if ncoord == 1 else list(y)
        neg = [(-c) % self.q for c in ys]
        out = bytearray(b"".join(self._fq(c) for c in xs + ys))
        if tuple(reversed(ys)) > tuple(reversed(neg)):     # y > -y, compared from the highest coefficient
            out[-1] |= 0x80
        return bytes(out)

    def encode(self, tagged) -> bytes:
        kind, val = tagged
        if kind == "g1":
            return self._point(val, 1)
        if kind == "g2":
            return self._point(val, 2)
        if kind == "gt":
            from . import pairing as pr
            words = pr.to_words(val)                        # Montgomery limbs in tower order
            rinv = pow(1 << 384, -1, self.q)
            vals = [sum(int(words[6 * i + k]) << (64 * k) for k in range(6)) * rinv % self.q for i in range(12)]
            return b"".join(self._fq(v) for v in vals)
        raise ValueError(kind)

    def challenge(self, label: bytes, values) -> int:
        for v in values:
            self.sponge.absorb_bytes(self.encode(v))
        if label == b"U":
            return 0
        return self.sponge.squeeze_foreign(self.r)
