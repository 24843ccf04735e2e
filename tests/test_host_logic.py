"""CPU tests of the product's host-visible logic: the __host__ __device__ field / group / digit code compiled
with g++ (carry chains emulated), and the C-ABI library's exported symbols (no compute calls without a GPU)."""
import ctypes
import os
import random
import re
import subprocess

import numpy as np
import pytest

import helpers as h
from oracle import bls12_377 as o

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HC_DIR = os.path.join(ROOT, "tests", "host_check")


@pytest.fixture(scope="module")
def hc():
    so = os.path.join(HC_DIR, "libhostcheck.so")
    src = os.path.join(HC_DIR, "host_check.cpp")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-o", so, src])
    return ctypes.CDLL(so)


def P(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def w32(v, n):
    return np.array([(v >> (32 * i)) & 0xFFFFFFFF for i in range(n)], dtype=np.uint32)


def i32(a):
    return sum(int(x) << (32 * i) for i, x in enumerate(a))


def test_field_constants(hc):
    q = np.zeros(12, np.uint32); qo = q.copy(); q2 = q.copy()
    r = np.zeros(8, np.uint32); ro = r.copy(); r2 = r.copy()
    hc.hc_consts(P(q), P(qo), P(q2), P(r), P(ro), P(r2))
    assert i32(q) == o.Q and i32(qo) == o.FQ_R and i32(q2) == o.FQ_R * o.FQ_R % o.Q
    assert i32(r) == o.R_ORDER and i32(ro) == o.FR_R and i32(r2) == o.FR_R * o.FR_R % o.R_ORDER
    # the even/odd CIOS relies on p = 1 and -p^-1 = 0xffffffff (mod 2^32) for both fields
    for p in (o.Q, o.R_ORDER):
        assert p % (1 << 32) == 1 and (-pow(p, -1, 1 << 32)) % (1 << 32) == 0xFFFFFFFF


def test_fq_arithmetic(hc):
    rng = random.Random(1)
    edge = [0, 1, o.Q - 1, o.Q - 2, o.FQ_R, 1 << 376, 0xFFFFFFFF, 1 << 32, (1 << 64) - 1, o.Q >> 1]
    vals = edge + [rng.randrange(o.Q) for _ in range(1500)]
    out = np.zeros(12, np.uint32)
    for i, a in enumerate(vals):
        b = vals[(i * 7 + 3) % len(vals)]
        hc.hc_fq_mul(P(w32(a, 12)), P(w32(b, 12)), P(out)); assert i32(out) == a * b * o.FQ_RINV % o.Q
        hc.hc_fq_sqr(P(w32(a, 12)), P(out)); assert i32(out) == a * a * o.FQ_RINV % o.Q
        hc.hc_fq_add(P(w32(a, 12)), P(w32(b, 12)), P(out)); assert i32(out) == (a + b) % o.Q
        hc.hc_fq_sub(P(w32(a, 12)), P(w32(b, 12)), P(out)); assert i32(out) == (a - b) % o.Q
        hc.hc_fq_neg(P(w32(a, 12)), P(out)); assert i32(out) == (-a) % o.Q
    out2 = np.zeros(12, np.uint32)
    for a in edge[1:] + vals[20:60] + [1, 2, o.Q - 1, o.FQ_R, (o.Q + 1) // 2]:
        # binary extended Euclid (fq_inv) against its definition and against the Fermat ladder it replaced
        hc.hc_fq_inv(P(w32(a, 12)), P(out)); assert i32(out) * a % o.Q == o.FQ_R * o.FQ_R % o.Q
        hc.hc_fq_inv_fermat(P(w32(a, 12)), P(out2)); assert np.array_equal(out, out2)
    hc.hc_fq_inv(P(w32(0, 12)), P(out)); assert i32(out) == 0


def test_fr_arithmetic(hc):
    rng = random.Random(2)
    out = np.zeros(8, np.uint32)
    for _ in range(1500):
        a = rng.randrange(o.R_ORDER); b = rng.randrange(o.R_ORDER)
        hc.hc_fr_mul(P(w32(a, 8)), P(w32(b, 8)), P(out)); assert i32(out) == a * b * o.FR_RINV % o.R_ORDER
        hc.hc_fr_add(P(w32(a, 8)), P(w32(b, 8)), P(out)); assert i32(out) == (a + b) % o.R_ORDER
        hc.hc_fr_to_canonical(P(w32(a, 8)), P(out)); assert i32(out) == a * o.FR_RINV % o.R_ORDER


def test_group_law_exceptional_cases(hc):
    def aw(p): return h.pts_to_np([p])[0].view(np.uint32)
    def fa(a): return h.pt_from_np(a.view(np.uint64))
    pts, _ = o.rand_points(24, 3)
    T = pts[23]
    out = np.zeros(24, np.uint32)
    cases = [(pts[0], pts[1]), (pts[2], pts[2]), (pts[3], o.neg(pts[3])), (None, pts[4]), (pts[5], None), (None, None)]
    cases += [(pts[i], pts[i + 1]) for i in range(6, 16)]
    for p, q in cases:
        hc.hc_madd(P(aw(p)), P(aw(q)), P(out)); assert fa(out) == o.add(p, q)
        hc.hc_add(P(aw(p)), P(aw(q)), P(aw(T)), P(out)); assert fa(out) == o.add(p, q)
        hc.hc_dbl(P(aw(p)), P(aw(T)), P(out)); assert fa(out) == o.add(p, p)
    rng = random.Random(4)
    for k in [0, 1, 2, 3, o.R_ORDER - 1, rng.randrange(o.R_ORDER), rng.randrange(1 << 128)]:
        hc.hc_scalar_mul(P(aw(pts[7])), P(w32(k, 8)), P(out)); assert fa(out) == o.mul(k, pts[7])


def test_signed_digits(hc):
    rng = random.Random(5)
    d = np.zeros(128, np.int32)
    for c in range(3, 24):
        for s in [0, 1, o.R_ORDER - 1, 1 << 252, o.R_ORDER - 2] + [rng.randrange(o.R_ORDER) for _ in range(100)]:
            W = hc.hc_digits(P(w32(s, 8)), c, P(d))
            assert W == (254 + c - 1) // c
            assert sum(int(d[w]) << (c * w) for w in range(W)) == s
            assert all(-(1 << (c - 1)) <= int(d[w]) <= (1 << (c - 1)) for w in range(W))
            assert d[W - 1] >= 0


def _declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "testudo_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(tb200_[a-z0-9_]+)\s*\(", hdr)))


def test_abi_library_exports_every_declared_symbol():
    """The C-ABI .so loads on a CPU-only box and exports everything include/testudo_b200.h declares."""
    from testudo_b200 import _lib, build

    build.build()
    lib = _lib.load()
    declared = _declared_symbols()
    assert len(declared) >= 30
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
    assert set(_lib.SIGNATURES) == set(declared)


def test_no_cpu_fallback_without_gpu():
    """Without a CUDA device the engine refuses to work (no silent CPU path)."""
    import torch
    from testudo_b200 import _lib

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    lib = _lib.load()
    assert lib.tb200_init(-1) != 0
    assert b"CPU fallback" in lib.tb200_last_error() or b"CUDA" in lib.tb200_last_error()
    out = np.zeros(12, np.uint64)
    assert lib.tb200_msm_g1(None, None, 0, 0, P(out)) == -2  # TB200_E_STATE
    with pytest.raises(_lib.EngineError):
        _lib.check(lib.tb200_msm_g1(None, None, 0, 0, P(out)))


def test_product_never_imports_oracle():
    """The oracle is test infrastructure: nothing under testudo_b200/ or include/ imports, includes or loads it."""
    bad = re.compile(r"(from\s+oracle|import\s+oracle|liboracle|#include[^\n]*oracle|oracle\.cpu|oracle/_build)")
    for top in ("testudo_b200", "include"):
        for dirpath, _, files in os.walk(os.path.join(ROOT, top)):
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp")):
                    txt = open(os.path.join(dirpath, f), errors="ignore").read()
                    assert not bad.search(txt), f"{f} uses the oracle"
