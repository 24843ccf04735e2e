"""TEST INFRASTRUCTURE ONLY -- ctypes loader for oracle/_build/liboracle.so (cpu_msm.c).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may import this.
Arrays are numpy uint64 in the C-ABI layouts of include/testudo_b200.h.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liboracle.so")
_lib = None


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "cpu_msm.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return _SO


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        _lib = ctypes.CDLL(_SO)
        _lib.oracle_ark_window_bits.restype = ctypes.c_int
        _lib.oracle_ark_window_bits.argtypes = [ctypes.c_size_t]
        _lib.oracle_num_threads.restype = ctypes.c_int
    return _lib


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def _u64(a, shape=None):
    a = np.ascontiguousarray(a, dtype=np.uint64)
    if shape is not None:
        a = a.reshape(shape)
    return a


def msm_g1(bases, scalars, mont: bool = False):
    bases = _u64(bases, (-1, 12))
    scalars = _u64(scalars, (-1, 4))
    n = min(len(bases), len(scalars))
    out = np.zeros(12, dtype=np.uint64)
    lib().oracle_msm_g1(_p(bases), _p(scalars), ctypes.c_size_t(n), ctypes.c_int(int(mont)), _p(out))
    return out


def msm_g1_batch(bases, scalars, rows, cols, row_stride, col_stride, mont: bool = False):
    bases = _u64(bases, (-1, 12))
    scalars = _u64(scalars, (-1, 4))
    assert len(bases) >= cols
    out = np.zeros((rows, 12), dtype=np.uint64)
    lib().oracle_msm_g1_batch(_p(bases), _p(scalars), ctypes.c_size_t(rows), ctypes.c_size_t(cols),
                              ctypes.c_ssize_t(row_stride), ctypes.c_ssize_t(col_stride), ctypes.c_int(int(mont)),
                              _p(out))
    return out


def g1_mul(p, k):
    p = _u64(p, (12,)); k = _u64(k, (4,))
    out = np.zeros(12, dtype=np.uint64)
    lib().oracle_g1_mul(_p(p), _p(k), _p(out))
    return out


def g1_add(p, q):
    p = _u64(p, (12,)); q = _u64(q, (12,))
    out = np.zeros(12, dtype=np.uint64)
    lib().oracle_g1_add(_p(p), _p(q), _p(out))
    return out


def compress_g1(vec, split, scaler, mont: bool = False):
    v = _u64(vec, (-1, 12)).copy()
    k = _u64(scaler, (4,))
    lib().oracle_compress_g1(_p(v), ctypes.c_size_t(split), _p(k), ctypes.c_int(int(mont)))
    return v[:split].copy()


def fq_binop(name, a, b):
    a = _u64(a, (6,)); b = _u64(b, (6,))
    out = np.zeros(6, dtype=np.uint64)
    getattr(lib(), "oracle_fq_" + name)(_p(a), _p(b), _p(out))
    return out


def fq_inv(a):
    a = _u64(a, (6,))
    out = np.zeros(6, dtype=np.uint64)
    lib().oracle_fq_inv(_p(a), _p(out))
    return out


def fr_from_mont(a):
    a = _u64(a, (4,))
    out = np.zeros(4, dtype=np.uint64)
    lib().oracle_fr_from_mont(_p(a), _p(out))
    return out


def gen_points(start, step, n):
    start = _u64(start, (12,)); step = _u64(step, (12,))
    out = np.zeros((n, 12), dtype=np.uint64)
    lib().oracle_gen_points(_p(start), _p(step), ctypes.c_size_t(n), _p(out))
    return out


def ark_window_bits(n: int) -> int:
    return lib().oracle_ark_window_bits(n)


def num_threads() -> int:
    return lib().oracle_num_threads()


def set_num_threads(n: int) -> None:
    """omp_set_num_threads: torchrun exports OMP_NUM_THREADS=1, which would make the CPU arm a 1-core figure."""
    lib().oracle_set_num_threads(ctypes.c_int(int(n)))
