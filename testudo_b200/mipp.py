"""Mirror of the G1 work in `mipp.rs`: `multiexponentiation`, `compress`, and the prover loop of
`MippProof::prove` (src/mipp.rs:31-153, 354-394). G2 / GT work (compress of h, pairing products, commit_g2) and the
Poseidon transcript are out of scope -- challenges come from a callback.

The vectors stay on the GPU across rounds (tb200_mipp_g1_*): upload once, two points back per round.
"""
from __future__ import annotations

import ctypes
from dataclasses import dataclass, field
from typing import Callable, List, Tuple

import numpy as np

from . import _lib, curve, fr, msm


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(ctypes.c_void_p)


class InvalidIPVectorLength(Exception):
    """`Error::InvalidIPVectorLength` (src/mipp.rs:400-421)."""


def multiexponentiation(left, right) -> np.ndarray:
    """src/mipp.rs:385-394: Err(InvalidIPVectorLength) on a length mismatch, else msm_unchecked(left, right)."""
    l = np.ascontiguousarray(left, dtype=np.uint64).reshape(-1, 12)
    r = np.ascontiguousarray(right, dtype=np.uint64).reshape(-1, 4)
    if len(l) != len(r):
        raise InvalidIPVectorLength()
    return msm.msm_unchecked(l, r)


def compress(vec, split: int, scaler_mont) -> np.ndarray:
    """src/mipp.rs:354-367 for G1: vec[i] = vec[i] + vec[i + split]^scaler (affine), result truncated to `split`."""
    v = np.ascontiguousarray(vec, dtype=np.uint64).reshape(-1, 12).copy()
    k = np.ascontiguousarray(scaler_mont, dtype=np.uint64).reshape(4)
    assert len(v) >= 2 * split
    _lib.check(_lib.engine().tb200_compress_g1(_ptr(v), split, _ptr(k), _lib.SCALARS_MONT))
    return v[:split].copy()


@dataclass
class MippProofG1:
    """G1 fields of `MippProof<E>` (src/mipp.rs:22-28): comms_u and final_a (+ the folded y for cross-checks)."""
    comms_u: List[Tuple[np.ndarray, np.ndarray]] = field(default_factory=list)
    final_a: np.ndarray = None
    final_y: np.ndarray = None
    xs: List[int] = field(default_factory=list)
    xs_inv: List[int] = field(default_factory=list)

    @classmethod
    def prove(cls, challenge: Callable[[bytes, List[np.ndarray]], int], a, y_mont, U) -> "MippProofG1":
        """G1 part of src/mipp.rs:31-153. `challenge(label, points)` returns c_inv as an integer mod r after the
        reference would have appended `points` (comm_u_l, comm_u_r; comm_t_l/r are GT and out of scope)."""
        lib = _lib.engine()
        a = np.ascontiguousarray(a, dtype=np.uint64).reshape(-1, 12)
        y = np.ascontiguousarray(y_mont, dtype=np.uint64).reshape(-1, 4)
        if len(a) != len(y):
            raise InvalidIPVectorLength()
        out = cls()
        challenge(b"U", [np.asarray(U)])                                 # transcript.append(b"U", U), :56
        h = ctypes.c_void_p()
        _lib.check(lib.tb200_mipp_g1_begin(_ptr(a), _ptr(y), len(a), _lib.SCALARS_MONT, ctypes.byref(h)))
        try:
            while lib.tb200_mipp_g1_len(h) > 1:                          # :58
                ul = np.zeros(12, dtype=np.uint64)
                ur = np.zeros(12, dtype=np.uint64)
                _lib.check(lib.tb200_mipp_g1_cross(h, _ptr(ul), _ptr(ur)))   # :77-85
                c_inv = challenge(b"challenge_i", [ul, ur]) % fr.R       # :97-101
                c = fr.inverse(c_inv)                                    # :106
                cw = curve.scalars_to_words([c], mont=True)[0]
                ciw = curve.scalars_to_words([c_inv], mont=True)[0]
                _lib.check(lib.tb200_mipp_g1_fold(h, _ptr(cw), _ptr(ciw)))   # compress(m_a, c), compress_field(m_y, c_inv)
                out.comms_u.append((ul, ur))                             # :117
                out.xs.append(c)
                out.xs_inv.append(c_inv)
            fa = np.zeros((1, 12), dtype=np.uint64)
            fy = np.zeros((1, 4), dtype=np.uint64)
            _lib.check(lib.tb200_mipp_g1_read(h, _ptr(fa), _ptr(fy)))
            out.final_a, out.final_y = fa[0], fy[0]                      # :122
        finally:
            _lib.check(lib.tb200_mipp_g1_end(h))
        return out
