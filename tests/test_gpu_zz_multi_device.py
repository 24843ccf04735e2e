"""Several GPUs driven by ONE process through the C ABI (tb200_init_devices): a large single MSM sharded by
point range, row commitments by row range over the replicated SRS, `Polynomial::commit` (rows + t) with one partial Miller
product per GPU, a pairing product by pair range -- each against the C / pairing oracle. Skipped below two devices.
Named zz so that it runs last: it appends devices to the session's engine."""
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


@pytest.fixture(scope="module")
def multi(engine):
    import torch

    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least two CUDA devices")
    nd = min(n, 8)
    _lib.init_devices(list(range(nd)))
    assert _lib.device_count() == nd
    engine.tb200_set_shard_min(256)     # shard even small inputs so the oracle sizes exercise the multi-GPU path
    yield nd
    engine.tb200_set_shard_min(0)


def _srs(n, seed):
    pts, _ = o.rand_points(n, seed)
    return h.pts_to_np(pts)


@pytest.mark.parametrize("n", [4099, 1 << 16])
def test_single_msm_sharded_by_point_range(engine, oracle_c, multi, n):
    pts, _ = o.rand_points(256, 5)
    bases = np.tile(h.pts_to_np(pts), (n // 256 + 1, 1))[:n].copy()
    sc = h.np_rand_scalars(n, n + 1)
    out = np.zeros(12, dtype=np.uint64)
    _lib.check(engine.tb200_msm_g1(P(bases), P(sc), n, 0, P(out)))
    assert np.array_equal(out, oracle_c.msm_g1(bases, sc))


def test_single_msm_sharded_resident(engine, oracle_c, multi):
    """tb200_msm_g1_sharded_dev: every GPU holds its own slice (uneven, one empty)."""
    import torch

    nd = multi
    pts, _ = o.rand_points(128, 6)
    sizes = [3000 + 517 * i for i in range(nd)]
    sizes[-1] = 0
    bases = [np.tile(h.pts_to_np(pts), (s // 128 + 1, 1))[:s].copy() for s in sizes]
    scal = [h.np_rand_scalars(s, 40 + i) if s else np.zeros((0, 4), dtype=np.uint64) for i, s in enumerate(sizes)]
    keep = []
    bp = (ctypes.c_void_p * nd)()
    sp = (ctypes.c_void_p * nd)()
    for i in range(nd):
        if sizes[i]:
            tb = torch.from_numpy(bases[i].view(np.int64)).to(f"cuda:{i}")
            ts = torch.from_numpy(scal[i].view(np.int64)).to(f"cuda:{i}")
            keep += [tb, ts]
            bp[i], sp[i] = tb.data_ptr(), ts.data_ptr()
    for i in range(nd):
        torch.cuda.synchronize(i)
    ns = (ctypes.c_size_t * nd)(*sizes)
    out = np.zeros(12, dtype=np.uint64)
    _lib.check(engine.tb200_msm_g1_sharded_dev(ctypes.cast(bp, ctypes.c_void_p), ctypes.cast(sp, ctypes.c_void_p),
                                               ctypes.cast(ns, ctypes.c_void_p), 0, P(out)))
    exp = oracle_c.msm_g1(np.concatenate(bases), np.concatenate(scal))
    assert np.array_equal(out, exp)


@pytest.mark.parametrize("rows,cols", [(37, 64), (1024, 256)])
def test_rows_sharded_by_row_range(engine, oracle_c, multi, rows, cols):
    srs = _srs(cols, 900 + cols)
    z = h.np_rand_scalars(rows * cols, rows)
    hnd = ctypes.c_void_p()
    _lib.check(engine.tb200_srs_load(P(srs), cols, 0, ctypes.byref(hnd)))
    out = np.zeros((rows, 12), dtype=np.uint64)
    _lib.check(engine.tb200_msm_g1_batch(hnd, P(z), rows, cols, 1, rows, 0, P(out)))          # un-transposed Z
    assert np.array_equal(out, oracle_c.msm_g1_batch(srs, z, rows, cols, 1, rows))
    out[:] = 0
    _lib.check(engine.tb200_msm_g1_batch(hnd, P(z), rows, cols, cols, 1, 0, P(out)))          # contiguous rows
    assert np.array_equal(out, oracle_c.msm_g1_batch(srs, z, rows, cols, cols, 1))
    bufs = [np.ascontiguousarray(z.reshape(rows, cols, 4)[i]).copy() for i in range(rows)]
    ptrs = (ctypes.c_void_p * rows)(*[b.ctypes.data for b in bufs])
    out[:] = 0
    _lib.check(engine.tb200_msm_g1_batch_ptrs(hnd, ctypes.cast(ptrs, ctypes.c_void_p), rows, cols, 0, P(out)))
    assert np.array_equal(out, oracle_c.msm_g1_batch(srs, z, rows, cols, cols, 1))
    _lib.check(engine.tb200_srs_free(hnd))


def test_hyrax_blinded_sharded(engine, oracle_c, multi):
    L, R = 300, 32
    G = _srs(R, 41)
    hpt = _srs(1, 42)[0]
    z = h.np_rand_scalars(L * R, 43)
    blinds = h.np_rand_scalars(L, 44)
    gens = commitments.MultiCommitGens(G, hpt)
    got = commitments.commit_inner(h.scalars_to_np(h.np_scalars_to_ints(z), mont=True),
                                   h.scalars_to_np(h.np_scalars_to_ints(blinds), mont=True), gens)
    ext = np.concatenate([G, hpt.reshape(1, 12)])
    zz = np.concatenate([z.reshape(L, R, 4), blinds.reshape(L, 1, 4)], axis=1).reshape(-1, 4)
    assert np.array_equal(got, oracle_c.msm_g1_batch(ext, zz, L, R + 1, R + 1, 1))
    gens.close()


@pytest.mark.parametrize("nv", [6, 12])
def test_sqrt_pst_commit_sharded(engine, oracle_c, multi, nv):
    """rows by row range, t from one partial Miller product per GPU + all-gather + ONE final exponentiation."""
    from oracle import pairing as opr

    m_col, m_row = nv // 2, nv - nv // 2
    rows, cols = 1 << m_col, 1 << m_row
    srs = _srs(cols, 700 + nv)
    z = h.np_rand_scalars(1 << nv, 70 + nv)
    qs, _ = o2.rand_points(rows, 71 + nv)
    h_vec = np.array([o2.affine_to_words(q) for q in qs], dtype=np.uint64).reshape(-1, 24)
    hnd = ctypes.c_void_p()
    _lib.check(engine.tb200_srs_load(P(srs), cols, 0, ctypes.byref(hnd)))
    out = np.zeros((rows, 12), dtype=np.uint64)
    t = np.zeros(72, dtype=np.uint64)
    _lib.check(engine.tb200_sqrt_pst_commit_strided(hnd, P(z), rows, cols, 1, rows, 0, P(h_vec), P(out), P(t)))
    exp_rows = oracle_c.msm_g1_batch(srs, z, rows, cols, 1, rows)
    assert np.array_equal(out, exp_rows)
    if nv <= 6:
        assert opr.from_words(t) == opr.multi_pairing([h.pt_from_np(r) for r in exp_rows], qs)
    else:
        engine.tb200_set_shard_min(1 << 30)          # single-GPU pairing product of the same pairs
        try:
            single = pairing.multi_pairing(exp_rows, h_vec)
        finally:
            engine.tb200_set_shard_min(256)
        assert np.array_equal(t, single)
    _lib.check(engine.tb200_srs_free(hnd))


def test_multi_pairing_sharded_by_pair_range(engine, multi):
    from oracle import pairing as opr

    n = 256 * multi                                   # the library shards pairing products from 256 pairs per GPU
    ps, _ = o.rand_points(4, 81)
    qs, _ = o2.rand_points(4, 82)
    g1 = np.tile(h.pts_to_np(ps), (n // 4, 1))
    g2 = np.tile(np.array([o2.affine_to_words(q) for q in qs], dtype=np.uint64).reshape(-1, 24), (n // 4, 1))
    got = pairing.multi_pairing(g1, g2)
    # prod_i e(P_i, Q_i)^(n/4) = e(sum (n/4) P_i ...): bilinearity gives the oracle a 4-pair job
    exp = opr.multi_pairing([o.mul(n // 4, p) for p in ps], qs)
    assert opr.from_words(got) == exp


def test_polynomial_commit_picks_host_path(engine, oracle_c, multi):
    nv = 12
    srs = _srs(1 << (nv - nv // 2), 91)
    z = h.np_rand_scalars(1 << nv, 92)
    ck = sqrt_pst.CommitterKey.from_points(srs)
    poly = sqrt_pst.Polynomial.from_evaluations(h.scalars_to_np(h.np_scalars_to_ints(z), mont=True))
    assert not poly.resident
    rows, _ = poly.commit(ck)
    assert np.array_equal(rows, oracle_c.msm_g1_batch(srs, z, 1 << (nv // 2), 1 << (nv - nv // 2), 1, 1 << (nv // 2)))
    ck.close()
