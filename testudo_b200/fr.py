"""Host-side Fr scalar glue (Python integers) for the mirrors of the reference's *unchanged* CPU code:
challenge inversion (src/mipp.rs:106), chi products (src/sqrt_pst.rs:152-166), q = Z * chi (src/sqrt_pst.rs:81-101).
None of this is on the MSM path; the engine only ever receives finished scalar vectors."""
from __future__ import annotations

from typing import List, Sequence

import numpy as np

from . import curve

R = curve.R_ORDER
_RINV = pow(curve.FR_R, -1, R)


def from_mont_words(arr) -> List[int]:
    a = np.asarray(arr, dtype=np.uint64).reshape(-1, 4)
    return [curve.from_limbs64(row) * _RINV % R for row in a]


def to_mont_words(vals: Sequence[int]) -> np.ndarray:
    return curve.scalars_to_words(vals, mont=True)


def get_chi_i(b: Sequence[int], i: int) -> int:
    """src/sqrt_pst.rs:152-166: chi_i(b) = prod_j (i_j ? b_j : 1 - b_j), bits of i taken MSB first."""
    m = len(b)
    prod = 1
    for j in range(m):
        if (i >> (m - j - 1)) & 1:
            prod = prod * b[j] % R
        else:
            prod = prod * (1 - b[j]) % R
    return prod


def inverse(x: int) -> int:
    return pow(x % R, -1, R)
