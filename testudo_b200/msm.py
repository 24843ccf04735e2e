"""Mirror of ark-ec 0.4 `VariableBaseMSM` for `ark_bls12_377::G1Projective`.

The reference reaches it at src/sqrt_pst.rs:198, src/mipp.rs:385-394, src/commitments.rs:70-86,
src/nizk/bullet.rs:93-118,237-257, src/dense_mlpoly.rs:553-555. Same names, argument meaning and error
behaviour; values are numpy uint64 arrays in ark's in-memory layout:
    bases   [n, 12]  x[6] || y[6] limbs, Montgomery; all-zero row == identity
    scalars [n, 4]   Fr limbs -- Montgomery form for `msm` / `msm_unchecked` (they take `&[Fr]`),
                     canonical for `msm_bigint` (it takes `&[BigInt<4>]`)
Results are the canonical affine point (`.into_affine()` of what arkworks returns) as a [12] uint64 array.
"""
from __future__ import annotations

import ctypes
from typing import Tuple, Union

import numpy as np

from . import _lib


def _u64(a, cols: int) -> np.ndarray:
    a = np.ascontiguousarray(a, dtype=np.uint64)
    return a.reshape(-1, cols)


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(ctypes.c_void_p)


def msm_bigint(bases, bigints) -> np.ndarray:
    """`VariableBaseMSM::msm_bigint(bases, bigints)`: canonical scalars; truncates to min(len)."""
    b = _u64(bases, 12)
    s = _u64(bigints, 4)
    n = min(len(b), len(s))
    out = np.zeros(12, dtype=np.uint64)
    _lib.check(_lib.engine().tb200_msm_g1(_ptr(b), _ptr(s), n, 0, _ptr(out)))
    return out


def msm_unchecked(bases, scalars) -> np.ndarray:
    """`VariableBaseMSM::msm_unchecked(bases, scalars)`: Montgomery-form `Fr` scalars; silently truncates to
    min(len) like arkworks (the `into_bigint()` conversion runs on the GPU)."""
    b = _u64(bases, 12)
    s = _u64(scalars, 4)
    n = min(len(b), len(s))
    out = np.zeros(12, dtype=np.uint64)
    _lib.check(_lib.engine().tb200_msm_g1(_ptr(b), _ptr(s), n, _lib.SCALARS_MONT, _ptr(out)))
    return out


def msm(bases, scalars) -> Tuple[str, Union[np.ndarray, int]]:
    """`VariableBaseMSM::msm`: ("ok", point) iff lengths match, else ("err", min_len) -- Rust's Result<_, usize>."""
    b = _u64(bases, 12)
    s = _u64(scalars, 4)
    if len(b) != len(s):
        return ("err", min(len(b), len(s)))
    return ("ok", msm_unchecked(b, s))


def g1_sum(points) -> np.ndarray:
    """Sum of a handful of affine points (combining per-GPU partial MSM results)."""
    p = _u64(points, 12)
    out = np.zeros(12, dtype=np.uint64)
    _lib.check(_lib.engine().tb200_g1_sum(_ptr(p), len(p), _ptr(out)))
    return out


def msm_each(bases, bigints, per_row: int) -> np.ndarray:
    """`rows` independent MSMs of `per_row` (<= 8) points each in one launch (tb200_msm_g1_each): bases [rows * per_row, 12],
    canonical scalars [rows * per_row, 4] -> [rows, 12]. The verifier's `commitment - g*value` and
    `g_mask_random[i] - g*point[i]` (ark-poly-commit `MultilinearPC::check`)."""
    b = _u64(bases, 12)
    s = _u64(bigints, 4)
    if len(b) != len(s) or per_row <= 0 or len(b) % per_row:
        raise ValueError("bases and scalars must hold rows * per_row entries")
    rows = len(b) // per_row
    out = np.zeros((rows, 12), dtype=np.uint64)
    _lib.check(_lib.engine().tb200_msm_g1_each(_ptr(b), _ptr(s), rows, per_row, 0, _ptr(out)))
    return out


class PendingMsm:
    """An MSM in flight (tb200_msm_g1_begin): `wait()` returns the affine point. Keeps its host inputs alive."""

    def __init__(self, handle, keep):
        self._h, self._keep = handle, keep

    def wait(self) -> np.ndarray:
        out = np.zeros(12, dtype=np.uint64)
        if self._h is not None:
            h, self._h = self._h, None
            _lib.check(_lib.engine().tb200_msm_g1_end(h, _ptr(out)))
            self._keep = None
        return out

    def __del__(self):
        try:
            if self._h is not None:
                _lib.engine().tb200_msm_g1_end(self._h, None)
        except Exception:
            pass


def msm_unchecked_begin(bases, scalars) -> PendingMsm:
    """`msm_unchecked` started without waiting: it runs on a side pipeline of the library next to whatever the caller
    does until `.wait()` (the reference's `try_par!` / `rayon::join` of independent MSMs, src/macros.rs:1-17)."""
    b = _u64(bases, 12)
    s = _u64(scalars, 4)
    n = min(len(b), len(s))
    handle = ctypes.c_void_p()
    _lib.check(_lib.engine().tb200_msm_g1_begin(_ptr(b), _ptr(s), n, _lib.SCALARS_MONT, ctypes.byref(handle)))
    return PendingMsm(handle, (b, s))


def msm_rows(bases, bigints, row_lengths) -> np.ndarray:
    """A ragged batch of independent small MSMs (rows of 0 .. 1024 points) in one launch (tb200_msm_g1_rows): row i takes
    the next row_lengths[i] entries; canonical scalars. Returns [rows, 12]."""
    b = _u64(bases, 12)
    s = _u64(bigints, 4)
    lens = np.ascontiguousarray(row_lengths, dtype=np.uint64).reshape(-1)
    if len(b) != len(s) or int(lens.sum()) != len(b):
        raise ValueError("bases and scalars must hold sum(row_lengths) entries")
    out = np.zeros((len(lens), 12), dtype=np.uint64)
    if len(lens):
        _lib.check(_lib.engine().tb200_msm_g1_rows(_ptr(b), _ptr(s), _ptr(lens), len(lens), 0, _ptr(out)))
    return out


class RowBatch:
    """Independent small MSMs collected from several places and run as ONE launch: `add` returns the row's index in the
    array `run()` returns."""

    def __init__(self):
        self._b, self._s, self._len = [], [], []

    def add(self, bases, bigints) -> int:
        b = _u64(bases, 12)
        s = _u64(bigints, 4)
        if len(b) != len(s):
            raise ValueError("bases and scalars differ in length")
        self._b.append(b)
        self._s.append(s)
        self._len.append(len(b))
        return len(self._len) - 1

    def run(self) -> np.ndarray:
        if not self._len:
            return np.zeros((0, 12), dtype=np.uint64)
        return msm_rows(np.concatenate(self._b), np.concatenate(self._s), self._len)
