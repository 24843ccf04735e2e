"""Mirror of the group work in `mipp.rs`: `multiexponentiation`, `compress`, and the prover loop of
`MippProof::prove` (src/mipp.rs:31-153, 354-394), including -- when the G2 key `h` and the CRS levels are passed --
the G2 `compress` of the commitment key (:114), the cross pairing products `comms_t` (:87-94, SURVEY.md 8f rank 3:
tb200_mipp_pairing_cross over the device-resident a and h), the structured polynomial (:128-131, 159-180), `commit_g2`
(:133) and the `open_g1` proof (:144). Challenges come from a `challenge(label, values)` callback that receives what
the reference appends: `PoseidonTranscript("fq").as_challenge()` (testudo_b200/poseidon_transcript.py) is the reference's own
transcript. `MippProofG1.verify` mirrors `MippProof::verify` (src/mipp.rs:182-333).

The G1 vectors stay on the GPU across rounds (tb200_mipp_g1_*): upload once, two points back per round.
"""
from __future__ import annotations

import ctypes
from dataclasses import dataclass, field
from typing import Callable, List, Tuple

import numpy as np

from . import _lib, curve, fr, msm, msm_g2, multilinear_pc


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(ctypes.c_void_p)


class InvalidIPVectorLength(Exception):
    """`Error::InvalidIPVectorLength` (src/mipp.rs:400-421)."""


def multiexponentiation(left, right) -> np.ndarray:
    """src/mipp.rs:385-394: Err(InvalidIPVectorLength) on a length mismatch, else msm_unchecked(left, right)."""
    l = np.ascontiguousarray(left, dtype=np.uint64).reshape(-1, 12)
    r = np.ascontiguousarray(right, dtype=np.uint64).reshape(-1, 4)
    if len(l) != len(r):
        raise InvalidIPVectorLength()
    return msm.msm_unchecked(l, r)


def compress(vec, split: int, scaler_mont) -> np.ndarray:
    """src/mipp.rs:354-367 for G1: vec[i] = vec[i] + vec[i + split]^scaler (affine), result truncated to `split`."""
    v = np.ascontiguousarray(vec, dtype=np.uint64).reshape(-1, 12).copy()
    k = np.ascontiguousarray(scaler_mont, dtype=np.uint64).reshape(4)
    assert len(v) >= 2 * split
    _lib.check(_lib.engine().tb200_compress_g1(_ptr(v), split, _ptr(k), _lib.SCALARS_MONT))
    return v[:split].copy()


@dataclass
class MippProofG1:
    """G1 fields of `MippProof<E>` (src/mipp.rs:22-28): comms_u and final_a (+ the folded y for cross-checks)."""
    comms_u: List[Tuple[np.ndarray, np.ndarray]] = field(default_factory=list)
    comms_t: List[Tuple[np.ndarray, np.ndarray]] = field(default_factory=list)   # GT pairs ([72] each), when `h` was passed
    final_a: np.ndarray = None
    final_y: np.ndarray = None
    xs: List[int] = field(default_factory=list)
    xs_inv: List[int] = field(default_factory=list)
    final_h: np.ndarray = None          # [24] G2 affine, when the key `h` was passed
    pst_proof_h: np.ndarray = None      # [m, 12] `ProofG1.proofs`, when the CRS levels were passed
    rs: List[int] = field(default_factory=list)

    def to_bytes(self, compress: bool = True) -> bytes:
        """`MippProof::serialize_with_mode` (benches/pst.rs:70-72)"""
        from . import serialize
        return serialize.mipp_proof_bytes(self.comms_t, self.comms_u, self.final_a, self.final_h, self.pst_proof_h, compress)

    @classmethod
    def from_bytes(cls, data: bytes, compress: bool = True) -> "MippProofG1":
        """`MippProof::deserialize_with_mode`: the proof as a verifier receives it (prover-side scratch fields stay empty)"""
        from . import serialize
        t, u, fa, fh, ph = serialize.mipp_proof_from_bytes(data, compress)
        return cls(comms_u=u, comms_t=t, final_a=fa, final_h=fh, pst_proof_h=ph)

    @classmethod
    def prove(cls, challenge: Callable[[bytes, List[np.ndarray]], int], a, y_mont, U, h=None,
              powers_of_g_levels=None, on_round: Callable[[int], None] = None) -> "MippProofG1":
        """src/mipp.rs:31-153. `challenge(label, values)` returns the squeezed scalar as an integer mod r after the
        reference would have appended `values` (comm_u_l, comm_u_r, and -- when the G2 key is passed -- comm_t_l, comm_t_r).
        `h` = `ck.powers_of_h[odd]` ([n, 24]); `powers_of_g_levels[i]` = `ck.powers_of_g[off + i]` for `open_g1`.
        `on_round(remaining_length)` is called after every round's folds are enqueued (and once before the first round):
        the caller's hook for starting independent work at a chosen point of the loop."""
        lib = _lib.engine()
        a = np.ascontiguousarray(a, dtype=np.uint64).reshape(-1, 12)
        y = np.ascontiguousarray(y_mont, dtype=np.uint64).reshape(-1, 4)
        if len(a) != len(y):
            raise InvalidIPVectorLength()
        out = cls()
        m_h = None                                                        # device handle of the folded G2 key
        if h is not None:
            hk = np.ascontiguousarray(h, dtype=np.uint64).reshape(-1, 24)
            if len(hk) != len(a):
                raise InvalidIPVectorLength()
            m_h = ctypes.c_void_p()
            _lib.check(lib.tb200_mipp_g2_begin(_ptr(hk), len(hk), _lib.SCALARS_MONT, ctypes.byref(m_h)))
        challenge(b"U", [np.asarray(U)])                                 # transcript.append(b"U", U), :56
        h_key, h = h, ctypes.c_void_p()
        _lib.check(lib.tb200_mipp_g1_begin(_ptr(a), _ptr(y), len(a), _lib.SCALARS_MONT, ctypes.byref(h)))
        try:
            if on_round is not None:
                on_round(len(a))
            while lib.tb200_mipp_g1_len(h) > 1:                          # :58
                ul = np.zeros(12, dtype=np.uint64)
                ur = np.zeros(12, dtype=np.uint64)
                appended = [ul, ur]
                if m_h is not None:        # :77-94 in one call: the cross MSMs and pairings_product(a_l, h_r), (a_r, h_l)
                    tl = np.zeros(72, dtype=np.uint64)                   # run side by side on separate streams
                    tr = np.zeros(72, dtype=np.uint64)
                    _lib.check(lib.tb200_mipp_cross_all(h, m_h, _ptr(ul), _ptr(ur), _ptr(tl), _ptr(tr)))
                    out.comms_t.append((tl, tr))                         # :116
                    appended += [tl, tr]
                else:
                    _lib.check(lib.tb200_mipp_g1_cross(h, _ptr(ul), _ptr(ur)))   # :77-85
                c_inv = challenge(b"challenge_i", appended) % fr.R       # :97-101
                c = fr.inverse(c_inv)                                    # :106
                cw = curve.scalars_to_words([c], mont=True)[0]
                ciw = curve.scalars_to_words([c_inv], mont=True)[0]
                _lib.check(lib.tb200_mipp_g1_fold(h, _ptr(cw), _ptr(ciw)))   # compress(m_a, c), compress_field(m_y, c_inv)
                if m_h is not None:                                      # compress(&mut m_h, split, &c_inv), :114
                    _lib.check(lib.tb200_mipp_g2_fold(m_h, _ptr(ciw)))     # enqueued: overlaps the next G1 round
                out.comms_u.append((ul, ur))                             # :117
                out.xs.append(c)
                out.xs_inv.append(c_inv)
                if on_round is not None:
                    on_round(int(lib.tb200_mipp_g1_len(h)))
            fa = np.zeros((1, 12), dtype=np.uint64)
            fy = np.zeros((1, 4), dtype=np.uint64)
            _lib.check(lib.tb200_mipp_g1_read(h, _ptr(fa), _ptr(fy)))
            out.final_a, out.final_y = fa[0], fy[0]                      # :122
        except BaseException:
            if m_h is not None:
                lib.tb200_mipp_g2_end(m_h)
            raise
        finally:
            _lib.check(lib.tb200_mipp_g1_end(h))
        if m_h is not None:
            assert lib.tb200_mipp_g2_len(m_h) == 1                       # :121
            fh = np.zeros((1, 24), dtype=np.uint64)
            try:
                _lib.check(lib.tb200_mipp_g2_read(m_h, _ptr(fh)))
            finally:
                _lib.check(lib.tb200_mipp_g2_end(m_h))
            out.final_h = fh[0]
            # structured polynomial p_h with final_h = h^{p_h(t)} (:128-131); commit_g2 is the reference's
            # debug_assert cross-check (:133-134) -- executed, like there
            ev_w = polynomial_evaluations_words(out.xs_inv)
            # The reference computes commit_g2 (its debug_assert cross-check, :133-134) and then open_g1 (:144); neither
            # feeds the other, and the challenges rs do not depend on commit_g2: open_g1 is STARTED first (it runs on the
            # library's side streams) and commit_g2 runs next to it.
            pending = None
            if powers_of_g_levels is not None:
                m = len(out.xs_inv)
                out.rs = [challenge(b"random_point", []) % fr.R for _ in range(m)]          # :138-141
                rs_w = curve.scalars_to_words(out.rs, mont=True) if m else np.zeros((0, 4), dtype=np.uint64)
                pending = multilinear_pc.open_g1_begin(powers_of_g_levels, ev_w, rs_w)     # :144
            c_h = msm_g2.msm_unchecked(h_key, ev_w)
            assert np.array_equal(c_h, out.final_h), "debug_assert!(c.h_product == final_h) (src/mipp.rs:134)"
            if pending is not None:
                out.pst_proof_h = pending.wait()
        return out


    def verify_prepare(self, vk: "multilinear_pc.VerifierKey", challenge: Callable[[bytes, List[np.ndarray]], int],
                       point: List[int], U, T, batch: "msm.RowBatch"):
        """Everything of `MippProof::verify` (src/mipp.rs:182-333) up to the group work that can be batched: the
        transcript replay, the TC fold `T * prod comm_t_l^(c_inv) comm_t_r^(c)` (one tb200_gt_multi_pow); the UC fold
        `U + sum c_inv comm_u_l + c comm_u_r` against final_u and the G1 fold of check_2 are QUEUED on `batch` (rows of a
        ragged MSM batch the caller runs once, together with its own). Returns None when a challenge has no inverse (the
        reference panics), else `finish(points)` -> (check_u, tc, [final_t, check_2 left, check_2 right] operand lists)."""
        from . import pairing
        m = len(self.comms_u)
        if len(self.comms_t) != m or len(point) < m:
            raise ValueError("one (comm_u, comm_t) pair and one point coordinate per round")
        U = np.ascontiguousarray(U, dtype=np.uint64).reshape(12)
        T = np.ascontiguousarray(T, dtype=np.uint64).reshape(pairing.GT_WORDS)
        xs, xs_inv, final_y = [], [], 1
        challenge(b"U", [U])                                                     # :203
        for i, ((ul, ur), (tl, tr)) in enumerate(zip(self.comms_u, self.comms_t)):
            c_inv = challenge(b"challenge_i", [ul, ur, tl, tr]) % fr.R          # :212-216
            if c_inv == 0:
                return None                                                      # `c_inv.inverse().unwrap()`, :218
            xs.append(fr.inverse(c_inv))
            xs_inv.append(c_inv)
            final_y = final_y * (1 + c_inv * point[i] - point[i]) % fr.R        # :226
        # :240-276, the fold / reduce over MippTU seeded with (T, U): exponent 1 for the seeds
        t_bases = np.stack([T] + [t for tl, tr in self.comms_t for t in (tl, tr)])
        u_bases = np.stack([U] + [u for ul, ur in self.comms_u for u in (ul, ur)])
        exps = curve.scalars_to_words([1] + [e for c, ci in zip(xs, xs_inv) for e in (ci, c)])
        tc = pairing.gt_multi_pow(t_bases, exps)
        rs = [challenge(b"random_point", []) % fr.R for _ in range(m)]           # :281-285
        v = 1
        for i in range(m):
            v = v * (1 + rs[i] * xs_inv[m - i - 1] - rs[i]) % fr.R              # :294-297
        # uc == final_u (:316) with final_u = final_a * final_y (:310), as ONE row: uc - final_y final_a is the identity
        # (the group law is exact on any curve point, so this is the same predicate for one MSM less)
        u_all = np.concatenate([u_bases, np.asarray(self.final_a, dtype=np.uint64).reshape(1, 12)])
        e_all = np.concatenate([exps, curve.scalars_to_words([(-final_y) % fr.R])])
        u_row = batch.add(u_all, e_all)
        finish_h = multilinear_pc.check_2_prepare(vk, self.final_h, rs, v, self.pst_proof_h, batch)      # check_h, :307
        final_t = (np.asarray(self.final_a).reshape(1, 12), np.asarray(self.final_h).reshape(1, 24))     # :311

        def finish(points):
            check_u = not points[u_row].any()
            return check_u, tc, [final_t] + finish_h(points)
        return finish

    def verify(self, vk: "multilinear_pc.VerifierKey", challenge: Callable[[bytes, List[np.ndarray]], int],
               point: List[int], U, T) -> bool:
        """`MippProof::verify(vk, transcript, proof, point, U, T)` (src/mipp.rs:182-333) for this proof (comms_u, comms_t,
        final_a, final_h, pst_proof_h). `challenge` replays the transcript exactly as `prove` drove it; all group and
        pairing work runs on the GPU; the reference's asserts become a False result."""
        from . import pairing
        batch = msm.RowBatch()
        finish = self.verify_prepare(vk, challenge, point, U, T, batch)
        if finish is None:
            return False
        check_u, tc, products = finish(batch.run())
        final_t, left, right = pairing.multi_pairing_batch(products)
        check_t = bool(np.array_equal(tc, final_t))                              # :313
        check_h = bool(np.array_equal(left, right))                              # :307-308
        return check_h and check_t and check_u


def polynomial_evaluations_words(cs_inv: List[int]) -> np.ndarray:
    """The same evaluations as [2^m, 4] Montgomery words, computed on the device (tb200_fr_subset_products): the Python
    loop below plus the conversion of 2^13 integers costs ~12 ms of host time per proof, the kernel nothing."""
    m = len(cs_inv)
    b = curve.scalars_to_words(cs_inv, mont=True) if m else np.zeros((0, 4), dtype=np.uint64)
    out = np.zeros((1 << m, 4), dtype=np.uint64)
    _lib.check(_lib.engine().tb200_fr_subset_products(_ptr(b), m, _ptr(out)))
    return out


def polynomial_evaluations_from_transcript(cs_inv: List[int]) -> List[int]:
    """src/mipp.rs:159-180: evaluations over {0,1}^m of prod_i (1 - z_i + cs_inv[m - i - 1] z_i); bit j of the index
    (from the lsb) selects cs_inv[m - j - 1]."""
    m = len(cs_inv)
    evals = [1]
    for j in range(m):
        f = cs_inv[m - j - 1] % fr.R
        evals = evals + [e * f % fr.R for e in evals]
    return evals
