"""Shared test helpers: conversions between the Python oracle's integers and the C-ABI numpy layouts."""
from __future__ import annotations

import json
import os

import numpy as np

from oracle import bls12_377 as o

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def pts_to_np(points) -> np.ndarray:
    return np.array([o.affine_to_words(p) for p in points], dtype=np.uint64).reshape(-1, 12)


def pt_from_np(row) -> o.Affine:
    return o.affine_from_words([int(x) for x in np.asarray(row, dtype=np.uint64).reshape(12)])


def scalars_to_np(scalars, mont: bool = False) -> np.ndarray:
    vals = [o.fr_to_mont(s) if mont else s for s in scalars]
    return np.array([o.to_limbs64(v, 4) for v in vals], dtype=np.uint64).reshape(-1, 4)


def scalars_from_np(arr, mont: bool = False):
    out = []
    for row in np.asarray(arr, dtype=np.uint64).reshape(-1, 4):
        v = o.from_limbs64(int(x) for x in row)
        out.append(o.fr_from_mont(v) if mont else v)
    return out


def fq_to_np(vals) -> np.ndarray:
    return np.array([o.to_limbs64(v, 6) for v in vals], dtype=np.uint64).reshape(-1, 6)


def fq_from_np(arr):
    return [o.from_limbs64(int(x) for x in row) for row in np.asarray(arr, dtype=np.uint64).reshape(-1, 6)]


def pt_hex(p: o.Affine):
    return None if p is None else [hex(p[0]), hex(p[1])]


def pt_unhex(h) -> o.Affine:
    return None if h is None else (int(h[0], 16), int(h[1], 16))


def load_golden(name: str):
    with open(os.path.join(GOLDEN_DIR, name)) as f:
        return json.load(f)


def np_rand_scalars(n: int, seed: int) -> np.ndarray:
    """Uniform canonical scalars < r as [n,4] uint64, fast (numpy) -- rejection on the top limb."""
    rng = np.random.default_rng(seed)
    out = rng.integers(0, 1 << 64, size=(n, 4), dtype=np.uint64)
    top = np.uint64(o.R_ORDER >> 192)
    out[:, 3] %= top  # strictly below the top limb of r => value < r (slightly non-uniform, fine for tests)
    return out


def np_scalars_to_ints(arr: np.ndarray):
    return [o.from_limbs64(int(x) for x in row) for row in arr]


def edge_case_inputs():
    """(name, points, scalars) covering the exceptional cases SURVEY.md 8c lists."""
    pts, _ = o.rand_points(8, 1234)
    r = o.R_ORDER
    cases = [
        ("all_zero_scalars", pts[:4], [0, 0, 0, 0]),
        ("all_one", pts[:4], [1, 1, 1, 1]),
        ("r_minus_1", pts[:2], [r - 1, r - 1]),
        ("top_bit_252", pts[:3], [1 << 252, (1 << 252) + 12345, r - 2]),
        ("identity_bases", [None, pts[1], None, pts[3]], [5, 6, 7, 8]),
        ("all_identity", [None, None], [3, 4]),
        ("duplicate_bases", [pts[0]] * 6, [7, 7, 7, 9, 9, 1]),
        ("p_and_minus_p", [pts[2], o.neg(pts[2])], [1234567, 1234567]),
        ("cancel_to_inf_mixed", [pts[2], o.neg(pts[2]), pts[3], pts[3]], [5, 5, r - 3, 3]),
        ("single", pts[:1], [0xDEADBEEF]),
        ("single_zero", pts[:1], [0]),
        ("window_boundaries", pts[:6], [(1 << 15), (1 << 16) - 1, (1 << 16), (1 << 31) + (1 << 15), (1 << 128) - 1, (1 << 200) + (1 << 15)]),
    ]
    return cases
