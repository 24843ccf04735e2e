"""GPU tests of the C ABI's contract: argument errors, concurrent callers, limits and a full-size closed form."""
import ctypes
import threading

import numpy as np
import pytest

import helpers as h
from oracle import bls12_377 as o
from testudo_b200 import _lib, msm, synthetic

pytestmark = pytest.mark.gpu


def P(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def test_argument_errors_are_reported_not_fatal(engine):
    out = np.zeros(12, np.uint64)
    assert engine.tb200_msm_g1(None, None, 5, 0, P(out)) == -1                       # TB200_E_ARG: null inputs, n > 0
    assert b"null" in engine.tb200_last_error()
    assert engine.tb200_msm_g1(None, None, 0, 0, None) == -1                         # null output
    assert engine.tb200_msm_g1_dev(ctypes.c_void_p(8), ctypes.c_void_p(8), 4, 0, ctypes.c_void_p(8), None) == -1  # alignment
    hsrs = ctypes.c_void_p()
    assert engine.tb200_srs_load(None, 4, 0, ctypes.byref(hsrs)) == -1
    pts, _ = o.rand_points(4, 9)
    B = h.pts_to_np(pts)
    _lib.check(engine.tb200_srs_load(P(B), 4, 0, ctypes.byref(hsrs)))
    z = h.scalars_to_np([1, 2, 3, 4, 5, 6])
    o6 = np.zeros((2, 12), np.uint64)
    assert engine.tb200_msm_g1_batch(hsrs, P(z), 2, 3, 3, 1, 0, P(o6)) == -1          # cols != SRS size
    assert engine.tb200_msm_g1_batch(hsrs, P(z), 0, 4, 4, 1, 0, P(o6)) == 0           # zero rows is a no-op
    _lib.check(engine.tb200_srs_free(hsrs))
    hm = ctypes.c_void_p()
    y = h.scalars_to_np([1, 2, 3])
    assert engine.tb200_mipp_g1_begin(P(B), P(y), 3, 0, ctypes.byref(hm)) == -1       # not a power of two
    # the engine still works after the errors
    assert h.pt_from_np(msm.msm_bigint(B, h.scalars_to_np([1, 1, 1, 1]))) == o.msm_naive(pts, [1, 1, 1, 1])


def test_concurrent_callers_like_rayon_workers(engine, oracle_c):
    """The reference issues MSMs from many rayon workers at once (src/sqrt_pst.rs:121-125, src/macros.rs:1-17):
    the ABI must be re-entrant. 8 threads x 6 MSMs of different sizes, every result checked."""
    bases = oracle_c.gen_points(h.pts_to_np([o.mul(17, o.G)])[0], h.pts_to_np([o.mul(19, o.G)])[0], 3000)
    jobs = []
    for t in range(8):
        for k in range(6):
            n = 100 + 371 * ((t * 6 + k) % 7)
            sc = h.np_rand_scalars(n, 1000 + t * 10 + k)
            jobs.append((t, n, sc, oracle_c.msm_g1(bases[:n], sc)))
    errors = []

    def worker(tid):
        try:
            for (t, n, sc, exp) in jobs:
                if t == tid and not np.array_equal(msm.msm_bigint(bases[:n], sc), exp):
                    errors.append((tid, n))
        except Exception as e:  # noqa: BLE001
            errors.append((tid, repr(e)))

    threads = [threading.Thread(target=worker, args=(t,)) for t in range(8)]
    for th in threads:
        th.start()
    for th in threads:
        th.join()
    assert not errors, errors


def test_single_msm_2p22_closed_form_and_skew(engine):
    """2^22 points generated on the GPU with known discrete logs (testudo_b200/synthetic.py): uniform and R1CS-like
    (50% zeros, 25% ones) scalars against the closed form (sum s_k dlog_k) * G."""
    n = 1 << 22
    bases = synthetic.make_bases_dev(n, seed=11)
    import torch

    out = torch.zeros(12, dtype=torch.int64, device="cuda")
    for skew in (False, True):
        sc = synthetic.make_scalars_dev(n, seed=12, skew=skew)
        _lib.check(engine.tb200_msm_g1_dev(ctypes.c_void_p(bases.data_ptr()), ctypes.c_void_p(sc.data_ptr()), n, 0,
                                           ctypes.c_void_p(out.data_ptr()), None))
        _lib.check(engine.tb200_stream_sync())
        assert np.array_equal(out.cpu().numpy().view(np.uint64), synthetic.expected_msm(sc, n, seed=11)), skew


def test_host_facing_msm_chunked_path(engine):
    """tb200_msm_g1 with host buffers splits n >= 2^21 into point-range chunks that accumulate into one persistent
    bucket array (the upload of chunk k+1 overlaps the compute of chunk k). Ragged size, uniform and skewed scalars,
    closed-form check; the device-resident call on the same inputs must give the identical point."""
    import torch

    n = (1 << 21) + 12345
    bases = synthetic.make_bases_dev(n, seed=21)
    for skew in (False, True):
        sc = synthetic.make_scalars_dev(n, seed=22, skew=skew)
        B = bases.cpu().numpy().view(np.uint64)
        S = sc.cpu().numpy().view(np.uint64)
        got = msm.msm_bigint(B, S)
        assert np.array_equal(got, synthetic.expected_msm(sc, n, seed=21)), skew
        out = torch.zeros(12, dtype=torch.int64, device="cuda")
        _lib.check(engine.tb200_msm_g1_dev(ctypes.c_void_p(bases.data_ptr()), ctypes.c_void_p(sc.data_ptr()), n, 0,
                                           ctypes.c_void_p(out.data_ptr()), None))
        _lib.check(engine.tb200_stream_sync())
        assert np.array_equal(out.cpu().numpy().view(np.uint64), got), skew


def test_pairing_argument_errors_and_concurrent_callers(engine):
    """Error returns of the pairing entry points and thread-safety: pairing products issued from
    several host threads at once (the reference runs its two pairings_product calls as rayon tasks, src/mipp.rs:87-94)."""
    from oracle import bls12_377_g2 as o2
    from oracle import pairing as pr
    from testudo_b200 import pairing

    gt = np.zeros(72, np.uint64)
    assert engine.tb200_multi_pairing(None, None, 3, P(gt)) == -1
    assert engine.tb200_multi_pairing(None, None, 0, None) == -1
    assert engine.tb200_miller_product(None, None, 2, P(gt)) == -1
    assert engine.tb200_gt_product_final_exp(None, 2, P(gt)) == -1
    assert engine.tb200_gt_pow(None, None, 0, 0, None) == 0                                  # empty batch is a no-op
    assert engine.tb200_gt_pow(None, None, 1, 0, P(gt)) == -1
    pts, _ = o.rand_points(4, 81)
    qs, _ = o2.rand_points(4, 82)
    A = h.pts_to_np(pts)
    B = np.array([o2.affine_to_words(q) for q in qs], dtype=np.uint64)
    y = h.scalars_to_np([1, 2, 3, 4], mont=True)
    ha, hh = ctypes.c_void_p(), ctypes.c_void_p()
    _lib.check(engine.tb200_mipp_g1_begin(P(A), P(y), 4, 1, ctypes.byref(ha)))
    _lib.check(engine.tb200_mipp_g2_begin(P(B[:2].copy()), 2, 1, ctypes.byref(hh)))
    u = np.zeros(12, np.uint64)
    assert engine.tb200_mipp_cross_all(ha, hh, P(u), P(u), P(gt), P(gt)) == -1               # lengths differ
    assert b"differ" in engine.tb200_last_error()
    assert engine.tb200_mipp_pairing_cross(ha, hh, P(gt), P(gt)) == -1
    assert engine.tb200_mipp_cross_all(ha, None, P(u), P(u), P(gt), P(gt)) == -1
    engine.tb200_mipp_g1_end(ha)
    engine.tb200_mipp_g2_end(hh)
    exp = np.array(pr.to_words(pr.multi_pairing(pts, qs)), dtype=np.uint64)
    results, errors = [None] * 6, []

    def worker(k):
        try:
            results[k] = pairing.multi_pairing(A, B)
        except Exception as e:  # pragma: no cover
            errors.append(e)

    threads = [threading.Thread(target=worker, args=(k,)) for k in range(6)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors and all(np.array_equal(r, exp) for r in results)
