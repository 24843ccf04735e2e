"""GPU parity of the C++ host mirror (testudo_b200/host/testudo_b200.hpp): commit / open / MIPP / Pedersen driven
from C++ through the C ABI, compared value by value with the big-integer oracle following src/sqrt_pst.rs and
src/mipp.rs."""
import os
import struct
import subprocess

import numpy as np
import pytest

import helpers as h
from oracle import bls12_377 as o
from testudo_b200 import build, fr

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DRV_DIR = os.path.join(ROOT, "tests", "host_cpp")


@pytest.fixture(scope="module")
def driver():
    build.build()
    exe = os.path.join(DRV_DIR, "host_driver")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-o", exe, os.path.join(DRV_DIR, "host_driver.cpp"),
                           "-L" + build.LIB_DIR, "-ltestudo_b200", "-Wl,-rpath," + build.LIB_DIR])
    return exe


class FakeTranscript:
    def __init__(self):
        self.st = 0xCBF29CE484222325

    def absorb(self, b: bytes):
        for x in b:
            self.st ^= x
            self.st = (self.st * 0x100000001B3) & 0xFFFFFFFFFFFFFFFF

    def challenge(self, label: bytes, pts):
        self.absorb(label)
        for p in pts:
            self.absorb(np.asarray(p, dtype=np.uint64).tobytes())
        return ((self.st | 1) + ((self.st >> 7) << 64)) % o.R_ORDER


@pytest.mark.parametrize("nv", [5, 8])
def test_cpp_host_mirror_matches_reference_semantics(engine, driver, tmp_path, nv):
    m_col = nv // 2
    m_row = nv - m_col
    z = o.rand_scalars(1 << nv, 900 + nv)
    srs, _ = o.rand_points(1 << m_row, 910 + nv)
    point = o.rand_scalars(nv, 920 + nv)
    hpt = o.mul(424243, o.G)
    blind = o.rand_scalars(1, 930 + nv)[0]
    blob = struct.pack("<Q", nv) + h.scalars_to_np(z, mont=True).tobytes() + h.pts_to_np(srs).tobytes() + \
        h.scalars_to_np(point, mont=True).tobytes() + h.pts_to_np([hpt]).tobytes() + h.scalars_to_np([blind], mont=True).tobytes()
    fin, fout = tmp_path / "in.bin", tmp_path / "out.bin"
    fin.write_bytes(blob)
    subprocess.check_call([driver, str(fin), str(fout)])
    data = np.frombuffer(fout.read_bytes(), dtype=np.uint64)
    pos = 0

    def take_pts(k):
        nonlocal pos
        v = data[pos: pos + 12 * k].reshape(k, 12)
        pos += 12 * k
        return v

    def take_fr():
        nonlocal pos
        v = data[pos: pos + 4]
        pos += 4
        return h.scalars_from_np(v, mont=True)[0]

    def take_u64():
        nonlocal pos
        v = int(data[pos])
        pos += 1
        return v

    comm_list = take_pts(1 << m_col)
    rows = [o.msm_naive(srs, [z[(j << m_col) | i] for j in range(1 << m_row)]) for i in range(1 << m_col)]
    assert [h.pt_from_np(r) for r in comm_list] == rows                                    # src/sqrt_pst.rs:121-125
    b = point[m_row:]
    chis = [fr.get_chi_i(b, i) for i in range(1 << m_col)]
    q = [sum(z[(j << m_col) | i] * chis[i] for i in range(1 << m_col)) % o.R_ORDER for j in range(1 << m_row)]
    u = h.pt_from_np(take_pts(1)[0])
    comm_q = h.pt_from_np(take_pts(1)[0])
    assert u == o.msm_naive(rows, chis) == o.msm_naive(srs, q) == comm_q                    # :198, :205-206
    a = point[:m_row]
    assert take_fr() == sum(qj * fr.get_chi_i(a, j) for j, qj in enumerate(q)) % o.R_ORDER  # eval, :105-115
    # MIPP loop, src/mipp.rs:58-120
    tr = FakeTranscript()
    tr.challenge(b"U", [h.pts_to_np([u])[0]])
    m_a, m_y = list(rows), list(chis)
    assert take_u64() == m_col
    while len(m_a) > 1:
        split = len(m_a) // 2
        u_l = o.msm_naive(m_a[:split], m_y[split:])
        u_r = o.msm_naive(m_a[split:], m_y[:split])
        got = take_pts(2)
        assert h.pt_from_np(got[0]) == u_l and h.pt_from_np(got[1]) == u_r
        c_inv = tr.challenge(b"challenge_i", [got[0], got[1]])
        c = pow(c_inv, -1, o.R_ORDER)
        m_a = [o.add(m_a[i], o.mul(c, m_a[split + i])) for i in range(split)]
        m_y = [(m_y[i] + c_inv * m_y[split + i]) % o.R_ORDER for i in range(split)]
    assert h.pt_from_np(take_pts(1)[0]) == m_a[0]
    assert take_fr() == m_y[0]
    # PedersenCommit::commit_slice, src/commitments.rs:79-86
    assert h.pt_from_np(take_pts(1)[0]) == o.add(o.msm_naive(srs, z[: 1 << m_row]), o.mul(blind, hpt))
    assert take_u64() == (1 << m_row) - 1        # VariableBaseMSM::msm -> Err(min_len)
    assert take_u64() == 1                       # multiexponentiation -> Err(InvalidIPVectorLength)


def test_cpp_host_mirror_g2_and_pst_openings(engine, driver, tmp_path):
    """The C++ mirror's G2 MSM / compress and MultilinearPC::open / open_g1 wrappers against the big-integer oracles."""
    from oracle import bls12_377_g2 as o2
    from oracle import pst

    nv = 4
    t = o.rand_scalars(nv, 940)
    evals = o.rand_scalars(1 << nv, 941)
    point = o.rand_scalars(nv, 942)
    exps = [pst.eq_exponents(t[k:]) for k in range(nv)]
    h_lv = [[o2.mul(e, o2.G2) for e in lv] for lv in exps]
    g_lv = [[o.mul(e, o.G) for e in lv] for lv in exps]
    g2np = lambda pts: np.array([o2.affine_to_words(p) for p in pts], dtype=np.uint64)
    blob = struct.pack("<Q", nv) + h.scalars_to_np(evals, mont=True).tobytes() + h.scalars_to_np(point, mont=True).tobytes()
    blob += b"".join(g2np(lv).tobytes() for lv in h_lv) + b"".join(h.pts_to_np(lv).tobytes() for lv in g_lv)
    fin, fout = tmp_path / "in2.bin", tmp_path / "out2.bin"
    fin.write_bytes(blob)
    subprocess.check_call([driver, str(fin), str(fout), "g2"])
    data = np.frombuffer(fout.read_bytes(), dtype=np.uint64)
    p2 = data[: 24 * nv].reshape(nv, 24)
    p1 = data[24 * nv: 36 * nv].reshape(nv, 12)
    rest = data[36 * nv:]
    assert [o2.affine_from_words(r) for r in p2] == pst.open_proofs(evals, point, h_lv, o2.msm_naive)
    assert [h.pt_from_np(r) for r in p1] == pst.open_proofs(evals, point, g_lv, o.msm_naive)
    assert o2.affine_from_words(rest[:24]) == o2.msm_naive(h_lv[0], evals)
    split = (1 << nv) // 2
    comp = rest[24:].reshape(split, 24)
    assert [o2.affine_from_words(r) for r in comp] == [o2.add(h_lv[0][i], o2.mul(point[0], h_lv[0][split + i]))
                                                       for i in range(split)]


def test_cpp_host_mirror_pairing(engine, driver, tmp_path):
    """The C++ mirror's pairing wrappers (E::multi_pairing, pairings_product, GT pow) against the big-integer oracle."""
    from oracle import bls12_377_g2 as o2
    from oracle import pairing as pr

    n = 5
    ps, _ = o.rand_points(n, 951)
    qs, _ = o2.rand_points(n, 952)
    ps[2] = None
    e = o.rand_scalars(1, 953)[0]
    g2np = np.array([o2.affine_to_words(p) for p in qs], dtype=np.uint64)
    blob = struct.pack("<Q", n) + h.pts_to_np(ps).tobytes() + g2np.tobytes() + h.scalars_to_np([e], mont=True).tobytes()
    fin, fout = tmp_path / "in3.bin", tmp_path / "out3.bin"
    fin.write_bytes(blob)
    subprocess.check_call([driver, str(fin), str(fout), "pairing"])
    data = np.frombuffer(fout.read_bytes(), dtype=np.uint64)
    gts = [pr.from_words(data[72 * i: 72 * i + 72]) for i in range(4)]
    first = pr.pairing(ps[0], qs[0])
    assert gts[0] == pr.multi_pairing(ps, qs)
    assert gts[1] == first
    assert gts[2] == pr.f12_pow(first, e)
    assert gts[3] == pr.multi_pairing(ps[:-1], qs[:-1])
    assert int(data[288]) == 1
    more = [pr.from_words(data[289 + 72 * i: 289 + 72 * i + 72]) for i in range(3)]
    assert more[0] == pr.f12_pow(pr.f12_mul(gts[0], first), e)          # pairing::multi_pow (tb200_gt_multi_pow)
    assert more[1] == gts[0] and more[2] == first                       # pairing::multi_pairing_batch
