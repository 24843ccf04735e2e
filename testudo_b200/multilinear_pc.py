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
from dataclasses import dataclass
from typing import Sequence

import numpy as np

from . import _lib, curve, msm, msm_g2, pairing


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


# ---- verifier side: `MultilinearPC::check` (ark-poly-commit 0.4) and the fork's `check_2` ---------------------------------
@dataclass
class VerifierKey:
    """ark-poly-commit `VerifierKey<E>{nv, g, h, g_mask_random}` plus the fork's `h_mask_random` (SURVEY.md App. A.3;
    field usage at src/circuit_verifier.rs:188-232,258-281). `MultilinearPC::trim` yields it next to the committer key:
    g_mask_random[i] = t_i g, h_mask_random[i] = t_i h for the trapdoor t. Arrays in ark's in-memory layout."""
    nv: int
    g: np.ndarray                  # [12]
    h: np.ndarray                  # [24]
    g_mask_random: np.ndarray      # [nv, 12]
    h_mask_random: np.ndarray      # [nv, 24]

    def __post_init__(self):
        self.g = np.ascontiguousarray(self.g, dtype=np.uint64).reshape(12)
        self.h = np.ascontiguousarray(self.h, dtype=np.uint64).reshape(24)
        self.g_mask_random = np.ascontiguousarray(self.g_mask_random, dtype=np.uint64).reshape(-1, 12)
        self.h_mask_random = np.ascontiguousarray(self.h_mask_random, dtype=np.uint64).reshape(-1, 24)
        if len(self.g_mask_random) != self.nv or len(self.h_mask_random) != self.nv:
            raise ValueError("the verifier key holds one mask per variable")


_ONE = np.array([1, 0, 0, 0], dtype=np.uint64)


def _neg_words(values: Sequence[int]) -> np.ndarray:
    """canonical limbs of -v mod r (scalar arithmetic stays on the host, as in the reference)"""
    return curve.scalars_to_words([(-int(v)) % curve.R_ORDER for v in values])


def check_prepare(vk: VerifierKey, commitment, point: Sequence[int], value: int, proofs, batch: "msm.RowBatch"):
    """First half of `check`: queues its nv + 1 G1 values on `batch` as two-point rows ({C, g} x {1, -v};
    {g_mask_i, g} x {1, -z_i} -- ark builds them from a fixed-base table of g) and returns `finish(points)`, which turns
    the batch's results into the operand lists [(g1s, g2s) left, (g1s, g2s) right] of the two pairing products."""
    c = np.ascontiguousarray(commitment, dtype=np.uint64).reshape(12)
    pi = np.ascontiguousarray(proofs, dtype=np.uint64).reshape(-1, 24)
    nv = vk.nv
    if len(point) != nv or len(pi) != nv:
        raise ValueError("point and proof must have vk.nv entries")          # the reference indexes 0..vk.nv (panics)
    neg = _neg_words([value] + list(point))
    rows = []
    for i in range(nv + 1):
        rows.append(batch.add(np.stack([c if i == 0 else vk.g_mask_random[i - 1], vk.g]), np.stack([_ONE, neg[i]])))

    def finish(points):
        pts = points[rows]
        return [(pts[:1], vk.h.reshape(1, 24)), (pts[1:], pi)]
    return finish


def check_products(vk: VerifierKey, commitment, point: Sequence[int], value: int, proofs):
    """The two pairing products of `check` as operand lists [left, right]; the G1 values in ONE launch."""
    batch = msm.RowBatch()
    finish = check_prepare(vk, commitment, point, value, proofs, batch)
    return finish(batch.run())


def check(vk: VerifierKey, commitment, point: Sequence[int], value: int, proofs) -> bool:
    """`MultilinearPC::check(vk, commitment, point, value, proof)` (ark-poly-commit 0.4 multilinear_pc/mod.rs; called at
    src/sqrt_pst.rs:261; gadget form src/circuit_verifier.rs:244-312):

        e(C - v g, h) == prod_i e(g_mask_random[i] - point[i] g, proof_i)

    Both sides are pairing products on the GPU (one pass for the two), compared as field elements -- exactly the
    reference's `left == right`."""
    left, right = pairing.multi_pairing_batch(check_products(vk, commitment, point, value, proofs))
    return bool(np.array_equal(left, right))


def check_2_prepare(vk: VerifierKey, commitment_h, point: Sequence[int], value: int, proofs_g1, batch: "msm.RowBatch"):
    """First half of `check_2` (see there for the form of the right-hand side): queues the fold -sum_i z_i proof_i on
    `batch`, computes C_h - v h, and returns `finish(points)` -> [left, right] operand lists."""
    ch = np.ascontiguousarray(commitment_h, dtype=np.uint64).reshape(24)
    pi = np.ascontiguousarray(proofs_g1, dtype=np.uint64).reshape(-1, 12)
    m = len(point)
    off = vk.nv - m
    if off < 0 or len(pi) != m:
        raise ValueError("point longer than the key, or one proof per variable missing")
    row = batch.add(pi, _neg_words(point)) if m else None
    left_q = msm_g2.msm_bigint(np.stack([ch, vk.h]), np.stack([_ONE, _neg_words([value])[0]]))
    left = (vk.g.reshape(1, 12), left_q.reshape(1, 24))

    def finish(points):
        if m == 0:
            return [left, (np.zeros((0, 12), dtype=np.uint64), np.zeros((0, 24), dtype=np.uint64))]
        return [left, (np.concatenate([pi, points[row].reshape(1, 12)]),
                       np.concatenate([vk.h_mask_random[off:off + m], vk.h.reshape(1, 24)]))]
    return finish


def check_2_products(vk: VerifierKey, commitment_h, point: Sequence[int], value: int, proofs_g1):
    """The two pairing products of `check_2` as operand lists [left, right]."""
    batch = msm.RowBatch()
    finish = check_2_prepare(vk, commitment_h, point, value, proofs_g1, batch)
    return finish(batch.run())


def check_2(vk: VerifierKey, commitment_h, point: Sequence[int], value: int, proofs_g1) -> bool:
    """The fork's `MultilinearPC::check_2(vk, &CommitmentG2, point, value, &ProofG1)` (src/mipp.rs:307; gadget form
    src/circuit_verifier.rs:170-241), the mirror image of `check` for a G2 commitment with G1 proofs:

        e(g, C_h - v h) == prod_i e(proof_i, h_mask_random[off + i] - point[i] h),   off = vk.nv - len(point)

    Left: the G2 value is a two-point G2 MSM. Right: the G2 operands are combinations of the KEY's points (h and its
    masks, order r), so bilinearity in the second argument is exact for any first argument on the curve and
        prod_i e(proof_i, h_mask_i - z_i h) = prod_i e(proof_i, h_mask_i) * e(-sum_i z_i proof_i, h):
    the same GT element with ONE small G1 MSM instead of m scalar multiplications in G2 (~3 ms each). `check` cannot be
    rewritten this way: its G2 operands are the untrusted proof."""
    left, right = pairing.multi_pairing_batch(check_2_products(vk, commitment_h, point, value, proofs_g1))
    return bool(np.array_equal(left, right))
