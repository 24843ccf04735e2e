"""Mirror of `commitments.rs`: `MultiCommitGens` (as a plain container; the Poseidon-seeded generator derivation of
src/commitments.rs:17-39 is out of scope) and `PedersenCommit::{commit_scalar, commit_slice}` (src/commitments.rs:70-86),
plus the Hyrax row fan-out `DensePolynomial::commit_inner` (src/dense_mlpoly.rs:315-329) as one batched GPU call."""
from __future__ import annotations

import ctypes
from typing import Optional

import numpy as np

from . import _lib, msm


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(ctypes.c_void_p)


class MultiCommitGens:
    """`MultiCommitGens<G>{n, G, h}` (src/commitments.rs:10-15) with G resident on the GPU for batched commits."""

    def __init__(self, G, h):
        self.G = np.ascontiguousarray(G, dtype=np.uint64).reshape(-1, 12)
        self.h = np.ascontiguousarray(h, dtype=np.uint64).reshape(12)
        self.n = len(self.G)
        self._srs: Optional[ctypes.c_void_p] = None
        self._srs_blinded: Optional[ctypes.c_void_p] = None

    def srs(self):
        if self._srs is None:
            hnd = ctypes.c_void_p()
            _lib.check(_lib.engine().tb200_srs_load(_ptr(self.G), self.n, 0, ctypes.byref(hnd)))
            self._srs = hnd
        return self._srs

    def srs_blinded(self):
        """G with h appended as one extra column (tb200_srs_load_blinded): the blind rides along as a row's last scalar."""
        if self._srs_blinded is None:
            hnd = ctypes.c_void_p()
            _lib.check(_lib.engine().tb200_srs_load_blinded(_ptr(self.G), self.n, _ptr(self.h), 0, ctypes.byref(hnd)))
            self._srs_blinded = hnd
        return self._srs_blinded

    def close(self):
        for name in ("_srs", "_srs_blinded"):
            if getattr(self, name) is not None:
                _lib.check(_lib.engine().tb200_srs_free(getattr(self, name)))
                setattr(self, name, None)


class PedersenCommit:
    @staticmethod
    def commit_scalar(scalar_mont, blind_mont, gens_n: MultiCommitGens) -> np.ndarray:
        """src/commitments.rs:70-77: msm_unchecked(&[G[0], h], &[scalar, blind])."""
        assert gens_n.n == 1
        bases = np.stack([gens_n.G[0], gens_n.h])
        scalars = np.stack([np.asarray(scalar_mont, dtype=np.uint64).reshape(4),
                            np.asarray(blind_mont, dtype=np.uint64).reshape(4)])
        return msm.msm_unchecked(bases, scalars)

    @staticmethod
    def commit_slice(scalars_mont, blind_mont, gens_n: MultiCommitGens) -> np.ndarray:
        """src/commitments.rs:79-86: msm_unchecked(&gens.G, scalars) + h * blind -- one MSM over G || h."""
        s = np.ascontiguousarray(scalars_mont, dtype=np.uint64).reshape(-1, 4)
        assert len(s) == gens_n.n
        bases = np.concatenate([gens_n.G, gens_n.h.reshape(1, 12)])
        scalars = np.concatenate([s, np.asarray(blind_mont, dtype=np.uint64).reshape(1, 4)])
        return msm.msm_unchecked(bases, scalars)


def commit_inner(Z_mont, blinds_mont, gens: MultiCommitGens) -> np.ndarray:
    """`DensePolynomial::commit_inner` (src/dense_mlpoly.rs:315-329): C[i] = commit_slice(Z[R*i .. R*(i+1)], blinds[i]).
    The L_size row MSMs over the shared gens.G are ONE batched call, with or without blinds: `commit(gens, false)` passes
    zero blinds and takes the plain SRS; with blinds (src/dense_mlpoly.rs:349-377) h is the extra column
    of the SRS and blinds[i] the extra scalar of row i."""
    z = np.ascontiguousarray(Z_mont, dtype=np.uint64).reshape(-1, 4)
    blinds = np.ascontiguousarray(blinds_mont, dtype=np.uint64).reshape(-1, 4)
    L = len(blinds)
    R = len(z) // L
    assert L * R == len(z) and R == gens.n
    lib = _lib.engine()
    rows = np.zeros((L, 12), dtype=np.uint64)
    if not blinds.any():
        _lib.check(lib.tb200_msm_g1_batch(gens.srs(), _ptr(z), L, R, R, 1, _lib.SCALARS_MONT, _ptr(rows)))
    else:
        _lib.check(lib.tb200_msm_g1_batch_blinded(gens.srs_blinded(), _ptr(z), L, R, _ptr(blinds), _lib.SCALARS_MONT,
                                                  _ptr(rows)))
    return rows
