"""Synthetic inputs with known discrete logarithms, generated ON THE GPU by the engine.

bases:   P[i * nb + j] = A_i + B_j,  A_i = (a0 + i*sa) * G,  B_j = (b0 + j*sb) * G  ->  dlog = a0 + i*sa + b0 + j*sb
scalars: uniform 4 x u64 limbs with the top limb reduced below r's top limb (canonical, < r)
The closed form  MSM = (sum_k s_k * dlog_k mod r) * G  needs only O(n) integer work on the host (numpy column
sums of 32-bit half limbs) plus ONE scalar multiplication on the GPU, so full-size runs are checked exactly.
"""
from __future__ import annotations

import ctypes
from typing import Tuple

import numpy as np
import torch

from . import _lib, curve


def _p(t):
    return ctypes.c_void_p(t.data_ptr())


def _np_p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def dlog_params(seed: int) -> Tuple[int, int, int, int]:
    rng = np.random.default_rng(seed ^ 0x5125_0001)
    vals = [int.from_bytes(rng.bytes(31), "little") % curve.R_ORDER or 1 for _ in range(4)]
    return tuple(vals)  # a0, sa, b0, sb


def _multiples_of_g(lib, start: int, step: int, count: int) -> np.ndarray:
    ks = curve.scalars_to_words([(start + i * step) % curve.R_ORDER for i in range(count)])
    g = np.tile(curve.generator_words(), (count, 1))
    out = np.zeros((count, 12), dtype=np.uint64)
    _lib.check(lib.tb200_test_g1_mul(_np_p(g), _np_p(ks), count, _np_p(out)))
    return out


def split(n: int) -> Tuple[int, int]:
    nb = 1 << (max(n.bit_length() - 1, 0) // 2)
    na = (n + nb - 1) // nb
    return na, nb


def make_bases_dev(n: int, seed: int = 1, multiples_of_g=None) -> torch.Tensor:
    """[n, 12] int64 CUDA tensor of affine points with known dlogs. `multiples_of_g(start, step, count)` -> [count, 12]
    words lets a checker supply the 2 sqrt(n) generator multiples from outside the engine (tests pass the oracle's)."""
    lib = _lib.init()
    a0, sa, b0, sb = dlog_params(seed)
    na, nb = split(n)
    gen = multiples_of_g or (lambda start, step, count: _multiples_of_g(lib, start, step, count))
    A = torch.from_numpy(np.ascontiguousarray(gen(a0, sa, na), dtype=np.uint64).view(np.int64)).cuda()
    B = torch.from_numpy(np.ascontiguousarray(gen(b0, sb, nb), dtype=np.uint64).view(np.int64)).cuda()
    out = torch.empty((na * nb, 12), dtype=torch.int64, device="cuda")
    _lib.check(lib.tb200_g1_outer_sum_dev(_p(A), na, _p(B), nb, _p(out), None))
    torch.cuda.synchronize()
    return out[:n]


def make_scalars_dev(n: int, seed: int = 1, skew: bool = False) -> torch.Tensor:
    """[n, 4] int64 CUDA tensor of canonical scalars < r. skew=True: 50% zeros, 25% ones, 25% uniform."""
    g = torch.Generator(device="cuda").manual_seed(seed)
    s = torch.randint(-(2 ** 63), 2 ** 63 - 1, (n, 4), dtype=torch.int64, device="cuda", generator=g)
    top = torch.randint(0, curve.R_ORDER >> 192, (n,), dtype=torch.int64, device="cuda", generator=g)
    s[:, 3] = top
    if skew:
        kind = torch.randint(0, 4, (n,), device="cuda", generator=g)
        s[kind < 2] = 0
        one = torch.tensor([1, 0, 0, 0], dtype=torch.int64, device="cuda")
        s[kind == 2] = one
    return s


def expected_dlog(scalars, n: int, seed: int = 1) -> int:
    """sum_k s_k * dlog_k mod r for MSM(make_bases_dev(n, seed), scalars): the discrete log of the result. Host integer
    work only; a checker turns it into the expected point with ITS OWN scalar multiplication (tests: the oracle's)."""
    a0, sa, b0, sb = dlog_params(seed)
    na, nb = split(n)
    s = scalars[:n].cpu().numpy().view(np.uint64) if isinstance(scalars, torch.Tensor) else np.asarray(scalars[:n]).view(np.uint64)
    pad = na * nb - n
    if pad:
        s = np.concatenate([s, np.zeros((pad, 4), dtype=np.uint64)])
    half = np.empty((na * nb, 8), dtype=np.uint64)  # 32-bit half limbs: column sums stay < 2^64
    half[:, 0::2] = s & np.uint64(0xFFFFFFFF)
    half[:, 1::2] = s >> np.uint64(32)
    half = half.reshape(na, nb, 8)
    row_sums = half.sum(axis=1, dtype=np.uint64) if nb < (1 << 31) else None  # [na, 8]
    col_sums = half.sum(axis=0, dtype=np.uint64)  # [nb, 8]

    def to_int(v):
        return sum(int(x) << (32 * k) for k, x in enumerate(v))

    total = 0
    for i in range(na):
        total += to_int(row_sums[i]) * (a0 + i * sa + b0)
    for j in range(nb):
        total += to_int(col_sums[j]) * (j * sb)
    return total % curve.R_ORDER


def expected_msm(scalars, n: int, seed: int = 1) -> np.ndarray:
    """Closed-form result of MSM(make_bases_dev(n, seed), scalars) as C-ABI words, with the one scalar multiplication on
    the GPU (self-check only: three engine paths agreeing; the parity tests use expected_dlog + the oracle)."""
    lib = _lib.init()
    k = curve.scalars_to_words([expected_dlog(scalars, n, seed)])
    out = np.zeros((1, 12), dtype=np.uint64)
    g = curve.generator_words().reshape(1, 12)
    _lib.check(lib.tb200_test_g1_mul(_np_p(g), _np_p(k), 1, _np_p(out)))
    return out[0]
