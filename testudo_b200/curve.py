"""BLS12-377 constants and limb-encoding helpers for the host side (no curve arithmetic happens here).

Constants as in ark-bls12-377 0.4 (`Cargo.toml:24` of the reference).
"""
from __future__ import annotations

from typing import Iterable, List

import numpy as np

Q = 0x01AE3A4617C510EAC63B05C06CA1493B1A22D9F300F5138F1EF3622FBA094800170B5D44300000008508C00000000001
R_ORDER = 0x12AB655E9A2CA55660B44D1E5C37B00159AA76FED00000010A11800000000001
GX = 0x008848DEFE740A67C8FC6225BF87FF5485951E2CAA9D41BB188282C8BD37CB5CD5481512FFCD394EEAB9B16EB21BE9EF
GY = 0x01914A69C5102EFF1F674F5D30AFEEC4BD7FB348CA3E52D96D182AD44FB82305C2FE3D3634A9591AFD82DE55559C8EA6
FQ_R = (1 << 384) % Q
FR_R = (1 << 256) % R_ORDER


def limbs64(v: int, n: int) -> List[int]:
    return [(v >> (64 * i)) & 0xFFFFFFFFFFFFFFFF for i in range(n)]


def from_limbs64(limbs: Iterable[int]) -> int:
    out = 0
    for i, l in enumerate(limbs):
        out |= int(l) << (64 * i)
    return out


def generator_words() -> np.ndarray:
    """The G1 generator in the C-ABI layout (x || y, Montgomery limbs)."""
    return np.array(limbs64(GX * FQ_R % Q, 6) + limbs64(GY * FQ_R % Q, 6), dtype=np.uint64)


def scalars_to_words(vals: Iterable[int], mont: bool = False) -> np.ndarray:
    rows = [limbs64((v % R_ORDER) * FR_R % R_ORDER if mont else v % R_ORDER, 4) for v in vals]
    return np.array(rows, dtype=np.uint64).reshape(-1, 4)
