"""Timing of the shared-SRS batched MSM (sqrt_pst commit row stage `comm_list`, src/sqrt_pst.rs:121-125) with stage
breakdown; verifies every row against the closed form g^{p_i(t)} for an SRS with known discrete logs."""
import ctypes
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from testudo_b200 import _lib, curve  # noqa: E402
from testudo_b200.synthetic import make_bases_dev, make_scalars_dev, dlog_params, split  # noqa: E402

lib = _lib.init()
for nv in [int(a) for a in sys.argv[1:]] or [20, 26]:
    m_col = nv // 2
    m_row = nv - m_col
    rows, cols = 1 << m_col, 1 << m_row
    srs_dev = make_bases_dev(cols, seed=nv)
    srs = srs_dev.cpu().numpy().view(np.uint64)
    h = ctypes.c_void_p()
    t0 = time.time()
    _lib.check(lib.tb200_srs_load(srs.ctypes.data_as(ctypes.c_void_p), cols, int(os.environ.get("TB_C", "0")), ctypes.byref(h)))
    t_srs = time.time() - t0
    z = make_scalars_dev(rows * cols, seed=nv + 1)
    out = torch.zeros((rows, 12), dtype=torch.int64, device="cuda")
    lib.tb200_set_profiling(1)
    lib.tb200_set_accumulate_mode(int(os.environ.get("TB_MODE", "0")))
    for rep in range(2):
        torch.cuda.synchronize()
        t0 = time.time()
        _lib.check(lib.tb200_msm_g1_batch_dev(h, ctypes.c_void_p(z.data_ptr()), rows, cols, 1, rows, 0,
                                              ctypes.c_void_p(out.data_ptr()), None))
        torch.cuda.synchronize()
        dt = time.time() - t0
    cc = ctypes.c_int(); W = ctypes.c_int(); K = ctypes.c_int(); M = ctypes.c_uint64(); B = ctypes.c_uint64()
    lib.tb200_last_geometry(ctypes.byref(cc), ctypes.byref(W), ctypes.byref(M), ctypes.byref(B), ctypes.byref(K))
    stages = {s: round(lib.tb200_stage_ms(s.encode()), 3) for s in
              ("digits", "scan", "scatter", "accumulate", "fixup", "reduce", "finalize", "total")}
    # closed form for a few rows: dlog(G_j) known -> row i commitment = (sum_j z[j*rows + i] * dlog_j) * G
    a0, sa, b0, sb = dlog_params(nv)
    na, nb = split(cols)
    dl = [(a0 + (j // nb) * sa + b0 + (j % nb) * sb) % curve.R_ORDER for j in range(cols)]
    zc = z.view(rows * cols, 4)
    ok = True
    for i in (0, 1, rows // 2 + 3, rows - 1):
        col = zc[i::rows].cpu().numpy().view(np.uint64)
        tot = sum(curve.from_limbs64(col[j]) * dl[j] for j in range(cols)) % curve.R_ORDER
        k = curve.scalars_to_words([tot])
        exp = np.zeros((1, 12), dtype=np.uint64)
        g = curve.generator_words().reshape(1, 12)
        _lib.check(lib.tb200_test_g1_mul(g.ctypes.data_as(ctypes.c_void_p), k.ctypes.data_as(ctypes.c_void_p), 1,
                                         exp.ctypes.data_as(ctypes.c_void_p)))
        ok = ok and bool(np.array_equal(out[i].cpu().numpy().view(np.uint64), exp[0]))
    print(json.dumps({"num_vars": nv, "rows": rows, "cols": cols, "verified_rows": ok, "c": cc.value, "W": W.value,
                      "K": K.value, "entries": M.value, "buckets": B.value, "srs_load_s": round(t_srs, 3),
                      "commit_ms": round(dt * 1e3, 2), "Mpairs_per_s": round(rows * cols / dt / 1e6, 2), "stages": stages}))
    # get_q on the device: q = Z * chis over the resident matrix, HBM-bound (32 B per element)
    chis = make_scalars_dev(rows, seed=nv + 2)
    qv = torch.zeros((cols, 4), dtype=torch.int64, device="cuda")
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    st = torch.cuda.Stream()
    torch.cuda.synchronize()
    for rep in range(3):
        e0.record(st)
        _lib.check(lib.tb200_fr_matvec_dev(ctypes.c_void_p(z.data_ptr()), cols, rows, ctypes.c_void_p(chis.data_ptr()),
                                           ctypes.c_void_p(qv.data_ptr()), ctypes.c_void_p(st.cuda_stream)))
        e1.record(st)
        torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print(json.dumps({"get_q_num_vars": nv, "ms": round(ms, 3), "GBps": round(rows * cols * 32 / ms / 1e6, 1)}))
    _lib.check(lib.tb200_srs_free(h))
    del z, out
    torch.cuda.empty_cache()
