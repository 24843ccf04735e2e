"""Mirror of the pairing calls on the reference's commitment path, ark-ec 0.4 `Pairing` for
`ark_bls12_377::Bls12_377`:

    E::multi_pairing(g1s, g2s).0      src/sqrt_pst.rs:143 (the IPP commitment t), src/mipp.rs:396-398 (pairings_product)
    E::pairing(p, q).0                src/mipp.rs:311 (verifier)
    TargetField::pow(bigint)          src/mipp.rs:258-261 (verifier); with the products of :260,:268-271 -> gt_multi_pow

Values are numpy uint64 arrays in ark's in-memory layout: G1 affine [12], G2 affine [24], GT = Fq12 [72] (tower order
c0.c0.c0 ... c1.c2.c1, Montgomery limbs). Everything runs on the GPU (tb200_multi_pairing / tb200_gt_pow); there is no
CPU path.
"""
from __future__ import annotations

import ctypes

import numpy as np

from . import _lib

GT_WORDS = 72


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(ctypes.c_void_p)


def multi_pairing(g1s, g2s) -> np.ndarray:
    """`E::multi_pairing(a, b).0`. ark zips the two iterators (the shorter one bounds the product) and skips pairs with
    an identity on either side."""
    p = np.ascontiguousarray(g1s, dtype=np.uint64).reshape(-1, 12)
    q = np.ascontiguousarray(g2s, dtype=np.uint64).reshape(-1, 24)
    n = min(len(p), len(q))
    out = np.zeros(GT_WORDS, dtype=np.uint64)
    _lib.check(_lib.engine().tb200_multi_pairing(_ptr(p), _ptr(q), n, _ptr(out)))
    return out


def multi_pairing_batch(products) -> np.ndarray:
    """Several independent `E::multi_pairing(g1s, g2s).0` in ONE pass of the pairing engine (tb200_multi_pairing_batch):
    `products` is a list of (g1s, g2s); shorter products are padded with identity pairs, which contribute 1 as in ark.
    Returns [len(products), 72]."""
    items = []
    for g1s, g2s in products:
        p = np.ascontiguousarray(g1s, dtype=np.uint64).reshape(-1, 12)
        q = np.ascontiguousarray(g2s, dtype=np.uint64).reshape(-1, 24)
        n = min(len(p), len(q))
        items.append((p[:n], q[:n]))
    k = len(items)
    width = max([len(p) for p, _ in items] + [1])
    g1 = np.zeros((k, width, 12), dtype=np.uint64)
    g2 = np.zeros((k, width, 24), dtype=np.uint64)
    for i, (p, q) in enumerate(items):
        g1[i, :len(p)] = p
        g2[i, :len(q)] = q
    out = np.zeros((k, GT_WORDS), dtype=np.uint64)
    if k:
        _lib.check(_lib.engine().tb200_multi_pairing_batch(_ptr(g1), _ptr(g2), k, width, _ptr(out)))
    return out


def pairings_product(gs, hs) -> np.ndarray:
    """src/mipp.rs:396-398."""
    return multi_pairing(gs, hs)


def pairing(p, q) -> np.ndarray:
    """`E::pairing(p, q).0`."""
    return multi_pairing(np.asarray(p).reshape(1, 12), np.asarray(q).reshape(1, 24))


def gt_pow(bases, exps, mont: bool = False) -> np.ndarray:
    """Element-wise `base.pow(exp)` in GT; exponents are Fr limbs (canonical `BigInt<4>` by default)."""
    b = np.ascontiguousarray(bases, dtype=np.uint64).reshape(-1, GT_WORDS)
    e = np.ascontiguousarray(exps, dtype=np.uint64).reshape(-1, 4)
    if len(b) != len(e):
        raise ValueError("bases and exponents differ in length")
    out = np.zeros_like(b)
    _lib.check(_lib.engine().tb200_gt_pow(_ptr(b), _ptr(e), len(b), _lib.SCALARS_MONT if mont else 0, _ptr(out)))
    return out


def gt_multi_pow(bases, exps, mont: bool = False) -> np.ndarray:
    """prod_i base[i].pow(exp[i]) in GT: the TC half of the verifier's fold / reduce over `MippTU`
    (src/mipp.rs:240-271) in one call; an empty product is 1."""
    b = np.ascontiguousarray(bases, dtype=np.uint64).reshape(-1, GT_WORDS)
    e = np.ascontiguousarray(exps, dtype=np.uint64).reshape(-1, 4)
    if len(b) != len(e):
        raise ValueError("bases and exponents differ in length")
    out = np.zeros(GT_WORDS, dtype=np.uint64)
    _lib.check(_lib.engine().tb200_gt_multi_pow(_ptr(b), _ptr(e), len(b), _lib.SCALARS_MONT if mont else 0, _ptr(out)))
    return out


def miller_product(g1s, g2s) -> np.ndarray:
    """This process's share of a sharded pairing product: the product of the Miller-loop values of the given pairs,
    WITHOUT the final exponentiation ([72] words; only meaningful as input of `final_exponentiation_of_product`)."""
    p = np.ascontiguousarray(g1s, dtype=np.uint64).reshape(-1, 12)
    q = np.ascontiguousarray(g2s, dtype=np.uint64).reshape(-1, 24)
    n = min(len(p), len(q))
    out = np.zeros(GT_WORDS, dtype=np.uint64)
    _lib.check(_lib.engine().tb200_miller_product(_ptr(p), _ptr(q), n, _ptr(out)))
    return out


def final_exponentiation_of_product(parts) -> np.ndarray:
    """Multiplies partial Miller products (one per rank) and applies the single final exponentiation."""
    f = np.ascontiguousarray(parts, dtype=np.uint64).reshape(-1, GT_WORDS)
    out = np.zeros(GT_WORDS, dtype=np.uint64)
    _lib.check(_lib.engine().tb200_gt_product_final_exp(_ptr(f), len(f), _ptr(out)))
    return out
