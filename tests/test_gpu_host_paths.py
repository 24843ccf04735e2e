"""GPU parity for the host-facing (sharding / chunked-upload) entry points added in round 2, vs the C oracle:
row batches from a strided matrix, from per-row heap buffers and with Hyrax blinds; `Polynomial::commit` in one call
(rows + t); point-range passes of a single MSM that exceeds the per-pass entry limit; stream-ordered workspace reuse."""
import ctypes

import numpy as np
import pytest

import helpers as h
from oracle import bls12_377 as o
from oracle import bls12_377_g2 as o2
from testudo_b200 import _lib, commitments, pairing, sqrt_pst

pytestmark = pytest.mark.gpu


def P(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def _srs(n, seed):
    pts, _ = o.rand_points(n, seed)
    return h.pts_to_np(pts)


def _load(engine, pts, window=0):
    hnd = ctypes.c_void_p()
    _lib.check(engine.tb200_srs_load(P(pts), len(pts), window, ctypes.byref(hnd)))
    return hnd


@pytest.mark.parametrize("rows,cols", [(3, 8), (40, 64), (700, 32), (1024, 1024)])
def test_batch_col_major_host(engine, oracle_c, rows, cols):
    """Un-transposed sqrt_pst matrix Z[(j << m_col) | i] from HOST memory (row_stride 1): >= 4 x SMs rows take the chunked
    upload path (cudaMemcpy2D of column ranges, overlapped with the compute)."""
    srs = _srs(cols, 1000 + cols)
    z = h.np_rand_scalars(rows * cols, rows * 7 + cols)
    z[5] = 0
    hnd = _load(engine, srs)
    out = np.zeros((rows, 12), dtype=np.uint64)
    _lib.check(engine.tb200_msm_g1_batch(hnd, P(z), rows, cols, 1, rows, 0, P(out)))
    exp = oracle_c.msm_g1_batch(srs, z, rows, cols, 1, rows)
    assert np.array_equal(out, exp)
    _lib.check(engine.tb200_srs_free(hnd))


@pytest.mark.parametrize("rows,cols,pad", [(5, 16, 0), (700, 64, 0), (650, 32, 7)])
def test_batch_row_major_host(engine, oracle_c, rows, cols, pad):
    """Contiguous rows (Hyrax layout) with an optional row pitch larger than cols."""
    srs = _srs(cols, 2000 + cols)
    pitch = cols + pad
    z = h.np_rand_scalars(rows * pitch, rows + cols + pad)
    hnd = _load(engine, srs)
    out = np.zeros((rows, 12), dtype=np.uint64)
    _lib.check(engine.tb200_msm_g1_batch(hnd, P(z), rows, cols, pitch, 1, 0, P(out)))
    exp = oracle_c.msm_g1_batch(srs, z, rows, cols, pitch, 1)
    assert np.array_equal(out, exp)
    _lib.check(engine.tb200_srs_free(hnd))


def test_batch_arbitrary_strides_host(engine, oracle_c):
    rows, cols = 6, 16
    srs = _srs(cols, 31)
    z = h.np_rand_scalars(3 * rows * 2 * cols, 5)
    hnd = _load(engine, srs)
    out = np.zeros((rows, 12), dtype=np.uint64)
    _lib.check(engine.tb200_msm_g1_batch(hnd, P(z), rows, cols, 3, 3 * rows * 2, 0, P(out)))
    exp = oracle_c.msm_g1_batch(srs, z, rows, cols, 3, 3 * rows * 2)
    assert np.array_equal(out, exp)
    _lib.check(engine.tb200_srs_free(hnd))


@pytest.mark.parametrize("rows,cols,separate", [(9, 32, True), (700, 64, False), (700, 64, True)])
def test_batch_ptrs(engine, oracle_c, rows, cols, separate):
    """Rows as separate heap buffers (self.polys of `Polynomial::commit`, src/sqrt_pst.rs:48-62). `separate` = every
    row in its own allocation (one copy per row); otherwise rows that happen to be adjacent travel as one copy."""
    srs = _srs(cols, 3000 + cols)
    z = h.np_rand_scalars(rows * cols, 11 * rows).reshape(rows, cols, 4)
    bufs = [np.ascontiguousarray(z[i]).copy() for i in range(rows)] if separate else [z[i] for i in range(rows)]
    ptrs = (ctypes.c_void_p * rows)(*[b.ctypes.data for b in bufs])
    hnd = _load(engine, srs)
    out = np.zeros((rows, 12), dtype=np.uint64)
    _lib.check(engine.tb200_msm_g1_batch_ptrs(hnd, ctypes.cast(ptrs, ctypes.c_void_p), rows, cols, 0, P(out)))
    exp = oracle_c.msm_g1_batch(srs, z.reshape(-1, 4), rows, cols, cols, 1)
    assert np.array_equal(out, exp)
    _lib.check(engine.tb200_srs_free(hnd))


@pytest.mark.parametrize("L,R", [(4, 8), (64, 32), (700, 16)])
def test_hyrax_commit_inner_blinded(engine, oracle_c, L, R):
    """`DensePolynomial::commit_inner` WITH blinds (src/dense_mlpoly.rs:315-329, src/commitments.rs:80-86): h is the extra
    SRS column, the blind the extra scalar -- checked against MSM over G || h per row in the C oracle."""
    G = _srs(R, 4000 + R)
    hpt = _srs(1, 4999)[0]
    z = h.np_rand_scalars(L * R, L + R)
    blinds = h.np_rand_scalars(L, 77 + L)
    blinds[1] = 0
    z_m = h.scalars_to_np(h.np_scalars_to_ints(z), mont=True)
    b_m = h.scalars_to_np(h.np_scalars_to_ints(blinds), mont=True)
    gens = commitments.MultiCommitGens(G, hpt)
    got = commitments.commit_inner(z_m, b_m, gens)
    ext = np.concatenate([G, hpt.reshape(1, 12)])
    zz = np.concatenate([z.reshape(L, R, 4), blinds.reshape(L, 1, 4)], axis=1).reshape(-1, 4)
    exp = oracle_c.msm_g1_batch(ext, zz, L, R + 1, R + 1, 1)
    assert np.array_equal(got, exp)
    # unblinded call on the same gens still takes the plain SRS
    got0 = commitments.commit_inner(z_m, np.zeros_like(b_m), gens)
    assert np.array_equal(got0, oracle_c.msm_g1_batch(G, z, L, R, R, 1))
    gens.close()


def test_blinded_srs_rejects_unblinded_call(engine):
    G = _srs(8, 5)
    hnd = ctypes.c_void_p()
    _lib.check(engine.tb200_srs_load_blinded(P(G), 8, P(G[0]), 0, ctypes.byref(hnd)))
    assert engine.tb200_srs_size(hnd) == 8
    z = h.np_rand_scalars(16, 3)
    out = np.zeros((2, 12), dtype=np.uint64)
    assert engine.tb200_msm_g1_batch(hnd, P(z), 2, 8, 8, 1, 0, P(out)) == -1
    plain = _load(engine, G)
    assert engine.tb200_msm_g1_batch_blinded(plain, P(z), 2, 8, P(z), 0, P(out)) == -1
    _lib.check(engine.tb200_srs_free(hnd))
    _lib.check(engine.tb200_srs_free(plain))


@pytest.mark.parametrize("nv", [6, 9, 12])
def test_sqrt_pst_commit_one_call(engine, oracle_c, nv):
    """tb200_sqrt_pst_commit[_strided]: rows == the C oracle's, t == the pairing oracle's definition on small sizes and
    == tb200_multi_pairing(rows, h_vec) (itself oracle-checked in test_gpu_pairing.py) on the larger one."""
    from oracle import pairing as opr

    m_col, m_row = nv // 2, nv - nv // 2
    rows, cols = 1 << m_col, 1 << m_row
    srs = _srs(cols, 600 + nv)
    z = h.np_rand_scalars(1 << nv, 60 + nv)
    qs, _ = o2.rand_points(rows, 61 + nv)
    h_vec = np.array([o2.affine_to_words(q) for q in qs], dtype=np.uint64).reshape(-1, 24)
    hnd = _load(engine, srs)
    out = np.zeros((rows, 12), dtype=np.uint64)
    t = np.zeros(72, dtype=np.uint64)
    _lib.check(engine.tb200_sqrt_pst_commit_strided(hnd, P(z), rows, cols, 1, rows, 0, P(h_vec), P(out), P(t)))
    exp_rows = oracle_c.msm_g1_batch(srs, z, rows, cols, 1, rows)
    assert np.array_equal(out, exp_rows)
    if nv <= 6:
        assert opr.from_words(t) == opr.multi_pairing([h.pt_from_np(r) for r in exp_rows], qs)
    assert np.array_equal(t, pairing.multi_pairing(exp_rows, h_vec))
    # the pointer-array form (self.polys): row i = Z[i :: 2^m_col]
    bufs = [np.ascontiguousarray(z[i::rows]) for i in range(rows)]
    ptrs = (ctypes.c_void_p * rows)(*[b.ctypes.data for b in bufs])
    out2 = np.zeros_like(out)
    t2 = np.zeros_like(t)
    _lib.check(engine.tb200_sqrt_pst_commit(hnd, ctypes.cast(ptrs, ctypes.c_void_p), rows, cols, 0, P(h_vec), P(out2), P(t2)))
    assert np.array_equal(out2, exp_rows) and np.array_equal(t2, t)
    _lib.check(engine.tb200_srs_free(hnd))


def test_polynomial_commit_host_matrix(engine, oracle_c):
    """The mirror's non-resident path (what several GPUs use): same rows as the resident path."""
    nv = 10
    srs = _srs(1 << (nv - nv // 2), 71)
    z = h.scalars_to_np(o.rand_scalars(1 << nv, 72), mont=True)
    ck = sqrt_pst.CommitterKey.from_points(srs)
    a, _ = sqrt_pst.Polynomial.from_evaluations(z, resident=True).commit(ck)
    b, _ = sqrt_pst.Polynomial.from_evaluations(z, resident=False).commit(ck)
    assert np.array_equal(a, b)
    ck.close()


@pytest.mark.parametrize("n,limit", [(5000, 20000), (70001, 300000)])
def test_single_msm_point_range_passes(engine, oracle_c, n, limit):
    """A single MSM with more sorted entries than one pass may index (n x windows > limit; 2^32 in production) runs as
    point-range passes over ONE persistent bucket array; the limit is lowered so that a small input takes that path."""
    pts, _ = o.rand_points(64, 9)
    bases = np.tile(h.pts_to_np(pts), (n // 64 + 1, 1))[:n]
    sc = h.np_rand_scalars(n, n)
    exp = oracle_c.msm_g1(bases, sc)
    engine.tb200_set_pass_entries_max(limit)
    try:
        out = np.zeros(12, dtype=np.uint64)
        _lib.check(engine.tb200_msm_g1(P(bases), P(sc), n, 0, P(out)))
    finally:
        engine.tb200_set_pass_entries_max(0)
    assert np.array_equal(out, exp)
    out2 = np.zeros(12, dtype=np.uint64)
    _lib.check(engine.tb200_msm_g1(P(bases), P(sc), n, 0, P(out2)))
    assert np.array_equal(out2, exp)


def test_dev_calls_on_two_streams_share_the_workspace(engine, oracle_c):
    """Two tb200_msm_g1_dev calls on different caller streams use the same scratch arena: the second is ordered behind
    the first by the arena's last-use event (results were silently wrong when they overlapped)."""
    import torch

    n = 1 << 15
    pts, _ = o.rand_points(128, 21)
    bases = np.tile(h.pts_to_np(pts), (n // 128, 1))
    s1, s2 = h.np_rand_scalars(n, 1), h.np_rand_scalars(n, 2)
    d_b = torch.from_numpy(bases.view(np.int64)).cuda()
    d_s1 = torch.from_numpy(s1.view(np.int64)).cuda()
    d_s2 = torch.from_numpy(s2.view(np.int64)).cuda()
    d_o = torch.zeros((2, 12), dtype=torch.int64, device="cuda")
    st1, st2 = torch.cuda.Stream(), torch.cuda.Stream()
    torch.cuda.synchronize()
    for _ in range(3):
        _lib.check(engine.tb200_msm_g1_dev(d_b.data_ptr(), d_s1.data_ptr(), n, 0, d_o[0].data_ptr(), st1.cuda_stream))
        _lib.check(engine.tb200_msm_g1_dev(d_b.data_ptr(), d_s2.data_ptr(), n, 0, d_o[1].data_ptr(), st2.cuda_stream))
    torch.cuda.synchronize()
    got = d_o.cpu().numpy().view(np.uint64)
    assert np.array_equal(got[0], oracle_c.msm_g1(bases, s1))
    assert np.array_equal(got[1], oracle_c.msm_g1(bases, s2))
