"""Mirror of ark-poly-commit 0.4 `multilinear_pc::MultilinearPC::{open, open_g1}` (PST13 openings) above the C ABI.

The reference reaches them at src/sqrt_pst.rs:225 (`MultilinearPC::open(ck, &q, &a_rev)`: G2 proofs over
`ck.powers_of_h[*]`) and src/mipp.rs:144 (`open_g1`: the fork's mirror with G1 proofs over `ck.powers_of_g[*]`) --
SURVEY.md 2.3 rows X1 and M6. The quotient loop and one MSM per variable run on the GPU (tb200_pst_open_g1/g2); the
evaluations are uploaded once.

    level_bases  list of nv arrays: level i holds the 2^(nv - i) points of the CRS level used for variable i
                 ([*, 24] uint64 for G2, [*, 12] for G1; ark in-memory layout, Montgomery)
    evals        [2^nv, 4] Fr limbs, `poly.to_evaluations()` order
    point        [nv, 4] Fr limbs
Montgomery-form field elements by default (ark's `Fr`), canonical with mont=False. Returns the proofs as [nv, 24|12].
"""
from __future__ import annotations

import ctypes
from typing import Sequence

import numpy as np

from . import _lib


class PendingOpen:
    """An opening in flight (tb200_pst_open_g1/g2_begin): `wait()` returns the proofs."""

    def __init__(self, handle, nv: int, width: int):
        self._h, self._nv, self._width = handle, nv, width

    def wait(self) -> np.ndarray:
        out = np.zeros((self._nv, self._width), dtype=np.uint64)
        if self._h is not None:
            h, self._h = self._h, None
            _lib.check(_lib.engine().tb200_pst_open_end(h, out.ctypes.data_as(ctypes.c_void_p)))
        return out

    def __del__(self):
        try:
            if self._h is not None:
                _lib.engine().tb200_pst_open_end(self._h, None)
        except Exception:
            pass


def _open(level_bases: Sequence[np.ndarray], evals, point, mont: bool, g2: bool, asynchronous: bool = False):
    width = 24 if g2 else 12
    ev = np.ascontiguousarray(evals, dtype=np.uint64).reshape(-1, 4)
    pt = np.ascontiguousarray(point, dtype=np.uint64).reshape(-1, 4)
    nv = len(pt)
    if len(ev) != 1 << nv:
        raise ValueError("evaluations must have 2^len(point) entries")      # assert_eq!(polynomial.num_vars(), ..)
    if len(level_bases) < nv:
        raise ValueError("one CRS level per variable is required")
    levels = [np.ascontiguousarray(b, dtype=np.uint64).reshape(-1, width) for b in level_bases[:nv]]
    for i, b in enumerate(levels):
        if len(b) != 1 << (nv - i):
            raise ValueError(f"CRS level {i} must hold 2^{nv - i} points")
    ptrs = (ctypes.c_void_p * max(nv, 1))(*[b.ctypes.data for b in levels])
    if asynchronous:
        if nv == 0:
            return PendingOpen(None, 0, width)
        handle = ctypes.c_void_p()
        fn = _lib.engine().tb200_pst_open_g2_begin if g2 else _lib.engine().tb200_pst_open_g1_begin
        _lib.check(fn(ev.ctypes.data_as(ctypes.c_void_p), nv, pt.ctypes.data_as(ctypes.c_void_p),
                      ctypes.cast(ptrs, ctypes.c_void_p), _lib.SCALARS_MONT if mont else 0, ctypes.byref(handle)))
        return PendingOpen(handle, nv, width)
    out = np.zeros((nv, width), dtype=np.uint64)
    fn = _lib.engine().tb200_pst_open_g2 if g2 else _lib.engine().tb200_pst_open_g1
    _lib.check(fn(ev.ctypes.data_as(ctypes.c_void_p), nv, pt.ctypes.data_as(ctypes.c_void_p),
                  ctypes.cast(ptrs, ctypes.c_void_p), _lib.SCALARS_MONT if mont else 0,
                  out.ctypes.data_as(ctypes.c_void_p)))
    return out


def open(level_bases_h: Sequence[np.ndarray], evals, point, mont: bool = True) -> np.ndarray:  # noqa: A001
    """`MultilinearPC::open(ck, polynomial, point)` -> `Proof{proofs: Vec<G2Affine>}` as [nv, 24]."""
    return _open(level_bases_h, evals, point, mont, True)


def open_g1(level_bases_g: Sequence[np.ndarray], evals, point, mont: bool = True) -> np.ndarray:
    """Fork API `MultilinearPC::open_g1(ck, polynomial, point)` -> `ProofG1{proofs: Vec<G1Affine>}` as [nv, 12]."""
    return _open(level_bases_g, evals, point, mont, False)


def open_begin(level_bases_h: Sequence[np.ndarray], evals, point, mont: bool = True) -> PendingOpen:
    """`MultilinearPC::open` started without waiting: the caller overlaps it with other work and calls `.wait()`."""
    return _open(level_bases_h, evals, point, mont, True, asynchronous=True)


def open_g1_begin(level_bases_g: Sequence[np.ndarray], evals, point, mont: bool = True) -> PendingOpen:
    """`open_g1` started without waiting (tb200_pst_open_g1_begin)."""
    return _open(level_bases_g, evals, point, mont, False, asynchronous=True)
