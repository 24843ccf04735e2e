"""Mirror of ark-ec 0.4 `VariableBaseMSM` for `ark_bls12_377::G2Projective`.

The reference reaches it inside `MultilinearPC::open` (src/sqrt_pst.rs:225: the PST opening proofs are G2 MSMs over
`powers_of_h`) and through `commit_g2` / `compress` in MIPP (src/mipp.rs:114,133). Values are numpy uint64 arrays in
ark's in-memory layout:
    bases   [n, 24]  x.c0[6] || x.c1[6] || y.c0[6] || y.c1[6] limbs, Montgomery; all-zero row == identity
    scalars [n, 4]   Fr limbs -- Montgomery form for `msm_unchecked`, canonical for `msm_bigint`
Results are the canonical affine point as a [24] uint64 array. Same length rules as the G1 mirror (msm.py).
"""
from __future__ import annotations

import ctypes
from typing import Tuple, Union

import numpy as np

from . import _lib


def _u64(a, cols: int) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.uint64).reshape(-1, cols)


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(ctypes.c_void_p)


def _run(bases, scalars, flags: int) -> np.ndarray:
    b = _u64(bases, 24)
    s = _u64(scalars, 4)
    n = min(len(b), len(s))
    out = np.zeros(24, dtype=np.uint64)
    _lib.check(_lib.engine().tb200_msm_g2(_ptr(b), _ptr(s), n, flags, _ptr(out)))
    return out


def msm_bigint(bases, bigints) -> np.ndarray:
    """`VariableBaseMSM::msm_bigint(bases, bigints)`: canonical scalars; truncates to min(len)."""
    return _run(bases, bigints, 0)


def msm_unchecked(bases, scalars) -> np.ndarray:
    """`VariableBaseMSM::msm_unchecked(bases, scalars)`: Montgomery-form `Fr` scalars; truncates to min(len)."""
    return _run(bases, scalars, _lib.SCALARS_MONT)


def msm(bases, scalars) -> Tuple[str, Union[np.ndarray, int]]:
    """`VariableBaseMSM::msm`: ("ok", point) or ("err", min_len) on a length mismatch (arkworks' `Err(min_len)`)."""
    b = _u64(bases, 24)
    s = _u64(scalars, 4)
    if len(b) != len(s):
        return "err", min(len(b), len(s))
    return "ok", msm_unchecked(b, s)


def compress(vec, split: int, scaler, mont: bool = True) -> np.ndarray:
    """MIPP `compress` on a G2 vector (src/mipp.rs:354-367): returns vec[:split] + scaler * vec[split:2*split]."""
    v = _u64(vec, 24).copy()
    if len(v) < 2 * split:
        raise ValueError("vector shorter than 2 * split")
    k = np.ascontiguousarray(scaler, dtype=np.uint64).reshape(4)
    _lib.check(_lib.engine().tb200_compress_g2(_ptr(v), split, _ptr(k), _lib.SCALARS_MONT if mont else 0))
    return v[:split]
