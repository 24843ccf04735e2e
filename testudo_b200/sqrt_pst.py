"""Mirror of `sqrt_pst::Polynomial` (src/sqrt_pst.rs) for the G1 work of commit / open.

Same names and argument meaning as the reference; values are numpy uint64 arrays in ark's memory layout
(Fr: [.., 4] Montgomery limbs; G1Affine: [.., 12] Montgomery limbs, all-zero == identity).

What runs where:
  * `commit`  -- the row fan-out `polys.par_iter().map(|p| MultilinearPC::commit(ck, p))` (src/sqrt_pst.rs:121-125)
    is ONE batched GPU call over the SRS `ck.powers_of_g[0]` resident on the device. Z is kept un-transposed:
    row i of the reference's `polys` is the strided view Z[(j << m_col) | i] (src/sqrt_pst.rs:58), which the
    kernel reads coalesced, so `from_evaluations` costs nothing.
  * `open`    -- M2 `msm_unchecked(comms, chis)` (src/sqrt_pst.rs:198), M3 `MultilinearPC::commit(ck, q)` (:205)
    and the G1 part of `MippProof::prove` (:212) run on the GPU.
  * with the G2 side of the key (`powers_of_h`, `powers_of_g` levels) `open` also produces the PST proof
    `MultilinearPC::open(ck, &q, &a_rev)` (:218-225, G2 MSMs) and MIPP's `final_h` / `pst_proof_h`.
  * the pairing product `t = multi_pairing(comm_list, powers_of_h[odd])` (src/sqrt_pst.rs:131-144) and MIPP's
    `comms_t` run on the GPU too when ck carries the G2 side; with a G1-only key `commit`
    returns t = None.
  * `open` / `verify` take the Fiat-Shamir challenges from a `challenge(label, values)` callback that receives what the
    reference appends; `PoseidonTranscript("fq").as_challenge()` (poseidon_transcript.py) is the
    reference's own transcript (host code, as there).
  * `verify` -- `Polynomial::verify` (src/sqrt_pst.rs:232-267): the GT fold, the small G1 MSMs and the five pairing
    products of `MippProof::verify` and `MultilinearPC::check` on the GPU (tb200_gt_multi_pow, tb200_msm_g1_rows,
    tb200_multi_pairing_batch).
"""
from __future__ import annotations

import ctypes
from dataclasses import dataclass
from typing import Callable, List, Optional, Tuple

import numpy as np

from . import _lib, curve, fr, mipp, msm, multilinear_pc, pairing

PST_START_LEN = 4096     # `open`: the G2 opening of q starts when the MIPP vectors are this short (see `open`)


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(ctypes.c_void_p)


class CommitterKey:
    """G1 side of ark-poly-commit `CommitterKey<E>`: nv and powers_of_g[0] (2^nv points), loaded on the GPU once
    with its window tables (the `ck` argument of src/sqrt_pst.rs:117,168)."""

    def __init__(self, powers_of_g0: np.ndarray, handle, nv: int):
        self.powers_of_g0 = powers_of_g0
        self.nv = nv
        self._h = handle
        # optional rest of `CommitterKey<E>`: powers_of_g[k] / powers_of_h[k] with 2^(nv - k) points each
        self.powers_of_g: Optional[List[np.ndarray]] = None     # [*, 12] per level
        self.powers_of_h: Optional[List[np.ndarray]] = None     # [*, 24] per level

    def with_levels(self, powers_of_g: List[np.ndarray], powers_of_h: List[np.ndarray]) -> "CommitterKey":
        assert len(powers_of_g) == self.nv and len(powers_of_h) == self.nv
        assert all(len(l) == 1 << (self.nv - k) for k, l in enumerate(powers_of_g))
        assert all(len(l) == 1 << (self.nv - k) for k, l in enumerate(powers_of_h))
        self.powers_of_g, self.powers_of_h = list(powers_of_g), list(powers_of_h)
        return self

    @classmethod
    def from_points(cls, powers_of_g0, window_bits: int = 0) -> "CommitterKey":
        pts = np.ascontiguousarray(powers_of_g0, dtype=np.uint64).reshape(-1, 12)
        n = len(pts)
        assert n & (n - 1) == 0 and n > 0, "powers_of_g[0] has 2^nv points"
        h = ctypes.c_void_p()
        _lib.check(_lib.engine().tb200_srs_load(_ptr(pts), n, window_bits, ctypes.byref(h)))
        return cls(pts, h, n.bit_length() - 1)

    def close(self):
        if self._h is not None:
            _lib.check(_lib.engine().tb200_srs_free(self._h))
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def pc_commit(ck: CommitterKey, evals_mont: np.ndarray) -> np.ndarray:
    """`MultilinearPC::commit(ck, poly).g_product` for one polynomial given by its evaluations (Montgomery Fr)."""
    z = np.ascontiguousarray(evals_mont, dtype=np.uint64).reshape(-1, 4)
    out = np.zeros((1, 12), dtype=np.uint64)
    _lib.check(_lib.engine().tb200_msm_g1_batch(ck._h, _ptr(z), 1, len(z), len(z), 1, _lib.SCALARS_MONT, _ptr(out)))
    return out[0]


@dataclass
class OpenG1:
    """G1 outputs of `Polynomial::open`: U.g_product, the commitment to q (debug cross-check) and the MIPP G1 data."""
    u: np.ndarray
    comm_q: np.ndarray
    mipp: "mipp.MippProofG1"
    pst_proof: Optional[np.ndarray] = None   # `Proof{proofs: Vec<G2Affine>}` as [m_row, 24] when ck carries powers_of_h


class _DeviceBuffer:
    """cudaMalloc'ed copy of a numpy array (tb200_dev_*), freed with the owner."""

    def __init__(self, arr: np.ndarray):
        lib = _lib.engine()
        self.ptr = ctypes.c_void_p()
        self.nbytes = arr.nbytes
        _lib.check(lib.tb200_dev_alloc(max(arr.nbytes, 16), ctypes.byref(self.ptr)))
        if arr.nbytes:
            _lib.check(lib.tb200_dev_upload(self.ptr, _ptr(arr), arr.nbytes))

    def download(self, shape, dtype=np.uint64) -> np.ndarray:
        out = np.zeros(shape, dtype=dtype)
        _lib.check(_lib.engine().tb200_dev_download(_ptr(out), self.ptr, out.nbytes))
        return out

    def free(self):
        if self.ptr is not None:
            _lib.check(_lib.engine().tb200_dev_free(self.ptr))
            self.ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class Polynomial:
    def __init__(self, z: np.ndarray, m: int, odd: int, resident: Optional[bool] = None):
        self.Z = z            # [2^n, 4] Montgomery Fr, un-transposed (host copy)
        self.m = m            # m_col
        self.odd = odd
        self.q: Optional[np.ndarray] = None
        self.chis_b: Optional[np.ndarray] = None
        # The reference transposes Z here (src/sqrt_pst.rs:48-62). One GPU: we upload it instead, once, and commit and
        # get_q work on the resident matrix. Several GPUs (tb200_init_devices): commit hands the HOST matrix to the
        # library, which ships every GPU its row range (chunked, overlapped with the compute); the primary's copy for
        # get_q is uploaded when `open` first needs it.
        self.resident = (_lib.device_count() <= 1) if resident is None else resident
        self._dZ_buf: Optional[_DeviceBuffer] = _DeviceBuffer(z) if self.resident else None

    @property
    def _dZ(self) -> "_DeviceBuffer":
        if self._dZ_buf is None:
            self._dZ_buf = _DeviceBuffer(self.Z)
        return self._dZ_buf

    @classmethod
    def from_evaluations(cls, Z, resident: Optional[bool] = None) -> "Polynomial":
        """src/sqrt_pst.rs:32-75. len(Z) must be a power of two."""
        z = np.ascontiguousarray(Z, dtype=np.uint64).reshape(-1, 4)
        n = len(z)
        assert n > 0 and n & (n - 1) == 0
        num_vars = n.bit_length() - 1
        _lib.engine()
        return cls(z, num_vars // 2, num_vars % 2, resident)

    @property
    def m_row(self) -> int:
        return self.m + self.odd

    def row(self, i: int) -> np.ndarray:
        """polys[i].Z of the reference: Z[(j << m_col) | i] for j < 2^m_row."""
        return self.Z[i :: 1 << self.m]

    def commit(self, ck: CommitterKey) -> Tuple[np.ndarray, Optional[np.ndarray]]:
        """src/sqrt_pst.rs:117-149 -> (comm_list as [2^m_col, 12] g_products, t). t = multi_pairing(comm_list,
        ck.powers_of_h[odd]) as a [72] GT element (computed from the device-resident commitments) when ck carries
        powers_of_h, else None."""
        rows, cols = 1 << self.m, 1 << self.m_row
        assert cols == len(ck.powers_of_g0), "ck.powers_of_g[0] must have 2^m_row points"
        lib = _lib.engine()
        if not self.resident:
            return self._commit_host(ck)
        d_out = _DeviceBuffer(np.zeros((rows, 12), dtype=np.uint64))
        _lib.check(lib.tb200_msm_g1_batch_dev(ck._h, self._dZ.ptr, rows, cols, 1, rows, _lib.SCALARS_MONT, d_out.ptr, None))
        t = None
        if ck.powers_of_h is not None:                                  # src/sqrt_pst.rs:128-143
            h_vec = np.ascontiguousarray(ck.powers_of_h[self.odd], dtype=np.uint64).reshape(-1, 24)
            assert len(h_vec) == rows, "comm_list.len() == h_vec.len() (src/sqrt_pst.rs:129)"
            d_h = _DeviceBuffer(h_vec)
            d_t = _DeviceBuffer(np.zeros(pairing.GT_WORDS, dtype=np.uint64))
            _lib.check(lib.tb200_multi_pairing_dev(d_out.ptr, d_h.ptr, rows, d_t.ptr, None))
            _lib.check(lib.tb200_stream_sync())
            t = d_t.download((pairing.GT_WORDS,))
            d_h.free()
            d_t.free()
        _lib.check(lib.tb200_stream_sync())
        out = d_out.download((rows, 12))
        d_out.free()
        return out, t

    def _commit_host(self, ck: CommitterKey) -> Tuple[np.ndarray, Optional[np.ndarray]]:
        """commit from the HOST matrix through the sharding entry points: rows (and t) in one library call."""
        rows, cols = 1 << self.m, 1 << self.m_row
        lib = _lib.engine()
        out = np.zeros((rows, 12), dtype=np.uint64)
        if ck.powers_of_h is None:
            _lib.check(lib.tb200_msm_g1_batch(ck._h, _ptr(self.Z), rows, cols, 1, rows, _lib.SCALARS_MONT, _ptr(out)))
            return out, None
        h_vec = np.ascontiguousarray(ck.powers_of_h[self.odd], dtype=np.uint64).reshape(-1, 24)
        assert len(h_vec) == rows, "comm_list.len() == h_vec.len() (src/sqrt_pst.rs:129)"
        t = np.zeros(pairing.GT_WORDS, dtype=np.uint64)
        _lib.check(lib.tb200_sqrt_pst_commit_strided(ck._h, _ptr(self.Z), rows, cols, 1, rows, _lib.SCALARS_MONT,
                                                     _ptr(h_vec), _ptr(out), _ptr(t)))
        return out, t

    def get_q(self, point: List[int]) -> None:
        """src/sqrt_pst.rs:81-101 on the GPU: chis by k_fr_chis, q = Z * chis by k_fr_matvec.
        `point` as integers mod r."""
        assert len(point) == 2 * self.m + self.odd
        lib = _lib.engine()
        b = fr.to_mont_words(point[self.m + self.odd:]) if self.m else np.zeros((0, 4), dtype=np.uint64)
        pow_m = 1 << self.m
        chis = np.zeros((pow_m, 4), dtype=np.uint64)
        _lib.check(lib.tb200_fr_chis(_ptr(b), self.m, _ptr(chis)))
        nq = pow_m << self.odd
        d_chis = _DeviceBuffer(chis)
        d_q = _DeviceBuffer(np.zeros((nq, 4), dtype=np.uint64))
        _lib.check(lib.tb200_fr_matvec_dev(self._dZ.ptr, nq, pow_m, d_chis.ptr, d_q.ptr, None))   # Z stays resident
        _lib.check(lib.tb200_stream_sync())
        q = d_q.download((nq, 4))
        d_chis.free()
        d_q.free()
        self.q, self.chis_b = q, chis

    def get_q_host(self, point: List[int]) -> Tuple[np.ndarray, np.ndarray]:
        """The same with Python integers (the reference's CPU loop); cross-check for tests, small sizes only."""
        b = point[self.m + self.odd:]
        pow_m = 1 << self.m
        chis = [fr.get_chi_i(b, i) for i in range(pow_m)]
        zi = fr.from_mont_words(self.Z)
        zq = [sum(zi[(j << self.m) | i] * chis[i] for i in range(pow_m)) % fr.R for j in range(pow_m << self.odd)]
        return fr.to_mont_words(zq), fr.to_mont_words(chis)

    def eval(self, point: List[int]) -> int:
        """src/sqrt_pst.rs:105-115: q(a) = sum_j q[j] * chi_j(a), on the GPU."""
        a = point[: len(point) // 2 + self.odd]
        if self.q is None:
            self.get_q(point)
        lib = _lib.engine()
        chis_a = np.zeros((len(self.q), 4), dtype=np.uint64)
        _lib.check(lib.tb200_fr_chis(_ptr(fr.to_mont_words(a)), len(a), _ptr(chis_a)))
        out = np.zeros((1, 4), dtype=np.uint64)
        _lib.check(lib.tb200_fr_matvec(_ptr(self.q), 1, len(self.q), _ptr(chis_a), _ptr(out)))
        return fr.from_mont_words(out)[0]

    def open(self, challenge: Callable[[bytes, List[np.ndarray]], int], comm_list: np.ndarray, ck: CommitterKey,
             point: List[int], t: Optional[np.ndarray] = None) -> OpenG1:
        """src/sqrt_pst.rs:168-230 (`t` is accepted and unused, as `_T` in MippProof::prove, src/mipp.rs:38).
        `challenge(label, appended_values)` receives what the reference appends and returns the squeezed scalar
        (`PoseidonTranscript("fq").as_challenge()` for the reference's transcript)."""
        if self.q is None:
            self.get_q(point)
        assert self.chis_b is not None, "chis(b) should have been computed for q"
        assert len(self.chis_b) == len(comm_list)                      # src/sqrt_pst.rs:194
        c_u = msm.msm_unchecked(comm_list, self.chis_b)                # M2, src/sqrt_pst.rs:198
        h_vec = ck.powers_of_h[self.odd] if ck.powers_of_h is not None else None     # src/sqrt_pst.rs:207
        g_levels = ck.powers_of_g[self.odd:] if ck.powers_of_g is not None else None  # variable CRS: off = ck.nv - m
        # The reference computes the PST proof of q AFTER the MIPP proof (:218-225), but it depends only on q and the
        # point, not on the transcript: it runs on streams of its own NEXT TO the MIPP rounds. It is started once the
        # folded vectors are down to PST_START_LEN elements: the first round is throughput-bound (8192 Miller loops at
        # 2^26) and an opening started before it only shares its SMs, the later rounds are latency-bound and leave the GPU
        # idle (scripts/ab_open.py at 2^26: 73.6 ms started before the loop, 72.2 ms after the first round).
        pending_box = []

        def start_pst(remaining: int) -> None:
            if not pending_box and ck.powers_of_h is not None and remaining <= PST_START_LEN:
                a_rev = list(point[: self.m + self.odd])[::-1]                        # :218-222
                pending_box.append(multilinear_pc.open_begin(ck.powers_of_h, self.q,
                                                             curve.scalars_to_words(a_rev, mont=True)))  # :225
        # M3 only feeds the reference's debug_assert (:205-206) -- nothing on the prover's path waits for it: it is STARTED
        # here on a side pipeline (tb200_msm_g1_begin) and compared once the MIPP proof is done
        pending_q = msm.msm_unchecked_begin(ck.powers_of_g0, self.q)   # M3, src/sqrt_pst.rs:205
        proof = mipp.MippProofG1.prove(challenge, comm_list, self.chis_b, c_u, h_vec, g_levels, on_round=start_pst)   # :212-213
        start_pst(0)                                                   # vectors of length 1: no round ran
        comm_q = pending_q.wait()
        assert np.array_equal(c_u, comm_q), "debug_assert!(c_u == comm.g_product) (src/sqrt_pst.rs:206)"
        pst_proof = pending_box[0].wait() if pending_box else None
        return OpenG1(u=c_u, comm_q=comm_q, mipp=proof, pst_proof=pst_proof)

    @staticmethod
    def verify(challenge: Callable[[bytes, List[np.ndarray]], int], vk: "multilinear_pc.VerifierKey", U, point: List[int],
               v: int, pst_proof, mipp_proof: "mipp.MippProofG1", T) -> bool:
        """src/sqrt_pst.rs:232-267: `MippProof::verify` (U = A^y for the opening vector A of T) then
        `MultilinearPC::check(vk, U, a_rev, v, pst_proof)`. `U` is the `g_product` of the commitment `open` returned.
        The small G1 MSMs of the two checks run as ONE ragged batch (tb200_msm_g1_rows), their five pairing products --
        e(final_a, final_h), both sides of check_2, both sides of check -- in ONE pass of the pairing engine
        (tb200_multi_pairing_batch)."""
        n = len(point)
        odd = n % 2
        a = list(point[: n // 2 + odd])
        b = list(point[n // 2 + odd:])
        batch = msm.RowBatch()                                              # every small G1 MSM of the two checks: ONE launch
        finish_mipp = mipp_proof.verify_prepare(vk, challenge, b, U, T, batch)               # :249
        if finish_mipp is None:
            return False
        finish_check = multilinear_pc.check_prepare(vk, U, a[::-1], v, pst_proof, batch)     # :254-261
        points = batch.run()
        check_u, tc, products = finish_mipp(points)
        products += finish_check(points)
        final_t, l2, r2, l1, r1 = pairing.multi_pairing_batch(products)
        res_mipp = check_u and bool(np.array_equal(tc, final_t)) and bool(np.array_equal(l2, r2))   # :250
        return res_mipp and bool(np.array_equal(l1, r1))
