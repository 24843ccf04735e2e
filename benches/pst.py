#!/usr/bin/env python
"""Mirror of the reference's only live benchmark, benches/pst.rs (BASELINE configs[0]). Same loop (`for s in [4, 5, 20, 27]`,
benches/pst.rs:26 -- 27 is replaced by 26, the BASELINE size) and the same CSV columns (benches/pst.rs:13-21,93-96): power,
commit_time, opening_time, verification_time, proof_size, commiter_key_size -- times in ms.

  commit_time   = `Polynomial::commit` -> (comm_list, t): the row MSMs (src/sqrt_pst.rs:121-125) and the pairing product
                  (src/sqrt_pst.rs:131-144), Z resident on the device, results read back.
  opening_time  = `Polynomial::open` (src/sqrt_pst.rs:168-230): get_q, M2, M3, the MIPP proof (cross MSMs, cross pairing
                  products, G1/G2 folds, commit_g2, open_g1) and the G2 PST proof, over the reference's own Fiat-Shamir
                  transcript (`PoseidonTranscript<Fq>`, host sponge behind tb200_poseidon_*).
  proof_size    = compressed `Proof` + `MippProof` bytes, as benches/pst.rs:64-74 (testudo_b200/serialize.py).
  verification_time = `Polynomial::verify` (src/sqrt_pst.rs:232-267; benches/pst.rs:76-90) over a fresh transcript: the GT
                  fold (tb200_gt_multi_pow), the G1 folds, `check_2`, `check` and the pairings on the GPU; the bench
                  asserts the verdict like the reference (`assert!(res == true)`).
  commiter_key_size = compressed size of `CommitterKey{nv, powers_of_g, powers_of_h, g, h}` by formula.
Synthetic inputs like the reference (`F::rand(test_rng)`, `MultilinearPC::setup`): uniform scalars and a CRS
powers[k][x] = eq((t_k..), x) * generator with a random trapdoor, generated on the GPU.
"""
import csv
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from testudo_b200 import _lib, curve, multilinear_pc, serialize, sqrt_pst  # noqa: E402
from testudo_b200.poseidon_transcript import PoseidonTranscript  # noqa: E402
from testudo_b200.synthetic import make_scalars_dev  # noqa: E402

G2_GENERATOR = (
    233578398248691099356572568220835526895379068987715365179118596935057653620464273615301663571204657964920925606294,
    140913150380207355837477652521042157274541796891053068589147167627541651775299824604154852141315666357241556069118,
    63160294768292073209381361943935198908131692476676907196754037919244929611450776219210369229519898517858833747423,
    149157405641012693445398062341192467754805999074082136895788947234480009303640899064710353187729182149407503257491)


def crs_levels(lib, t, g2):
    """powers[k][x] = eq((t_k..t_{nv-1}), x) * generator (little-endian variables), ark-poly-commit `setup`."""
    import ctypes

    R = curve.R_ORDER
    if g2:
        gen = np.array(sum([curve.limbs64(c * curve.FQ_R % curve.Q, 6) for c in G2_GENERATOR], []), dtype=np.uint64).reshape(1, 24)
    else:
        gen = curve.generator_words().reshape(1, 12)
    fn = lib.tb200_test_g2_mul if g2 else lib.tb200_test_g1_mul
    out = []
    for k in range(len(t)):
        e = [1]
        for tj in t[k:]:
            e = [v * ((1 - tj) % R) % R for v in e] + [v * tj % R for v in e]
        pts = np.zeros((len(e), gen.shape[1]), dtype=np.uint64)
        g = np.ascontiguousarray(np.tile(gen, (len(e), 1)))
        kw = curve.scalars_to_words(e)
        _lib.check(fn(g.ctypes.data_as(ctypes.c_void_p), kw.ctypes.data_as(ctypes.c_void_p), len(e),
                      pts.ctypes.data_as(ctypes.c_void_p)))
        out.append(pts)
    return out


def verifier_key(lib, t) -> multilinear_pc.VerifierKey:
    """`MultilinearPC::trim` -> vk: g, h and the masks t_i g, t_i h of the trapdoor"""
    import ctypes

    def masks(gen, fn):
        g = np.ascontiguousarray(np.tile(gen, (len(t), 1)))
        out = np.zeros_like(g)
        kw = curve.scalars_to_words(list(t))
        _lib.check(fn(g.ctypes.data_as(ctypes.c_void_p), kw.ctypes.data_as(ctypes.c_void_p), len(t),
                      out.ctypes.data_as(ctypes.c_void_p)))
        return out

    g1 = curve.generator_words().reshape(1, 12)
    g2 = np.array(sum([curve.limbs64(c * curve.FQ_R % curve.Q, 6) for c in G2_GENERATOR], []), dtype=np.uint64).reshape(1, 24)
    return multilinear_pc.VerifierKey(nv=len(t), g=g1[0], h=g2[0], g_mask_random=masks(g1, lib.tb200_test_g1_mul),
                                      h_mask_random=masks(g2, lib.tb200_test_g2_mul))


def main():
    sizes = [int(a) for a in sys.argv[1:]] or [4, 5, 20, 26]
    lib = _lib.init()
    rows = []
    for s in sizes:
        m_row = s - s // 2
        rng = np.random.default_rng(1000 + s)
        z = make_scalars_dev(1 << s, seed=s).cpu().numpy().view(np.uint64)     # Montgomery-form Fr, like ark memory
        t = [int.from_bytes(rng.bytes(40), "little") % curve.R_ORDER for _ in range(m_row)]
        g_levels, h_levels = crs_levels(lib, t, False), crs_levels(lib, t, True)
        ck = sqrt_pst.CommitterKey.from_points(g_levels[0]).with_levels(g_levels, h_levels)   # setup + trim
        vk = verifier_key(lib, t)
        pl = sqrt_pst.Polynomial.from_evaluations(z)
        r = [int.from_bytes(rng.bytes(40), "little") % curve.R_ORDER for _ in range(s)]
        v = pl.eval(r)                                                           # benches/pst.rs:52

        def challenge():                                                         # PoseidonTranscript::new(&params), :58,:76
            return PoseidonTranscript("fq").as_challenge()

        pl.commit(ck)                                                            # warm-up (tables, arena)
        commit_ms = open_ms = float("inf")
        for _ in range(2):                                                       # wall-clock on a shared host: best of two
            t0 = time.perf_counter()
            comm_list, t_gt = pl.commit(ck)
            commit_ms = min(commit_ms, (time.perf_counter() - t0) * 1e3)
        pl.open(challenge(), comm_list, ck, r, t_gt)                             # warm-up
        for _ in range(2):
            pl.q = None
            tr = challenge()
            t0 = time.perf_counter()
            opened = pl.open(tr, comm_list, ck, r, t_gt)
            open_ms = min(open_ms, (time.perf_counter() - t0) * 1e3)
        verify_ms = float("inf")
        for _ in range(3):                                                       # first pass = warm-up
            tr = challenge()
            t0 = time.perf_counter()
            res = sqrt_pst.Polynomial.verify(tr, vk, opened.u, r, v, opened.pst_proof, opened.mipp, t_gt)
            verify_ms = min(verify_ms, (time.perf_counter() - t0) * 1e3)
            assert res is True                                                   # benches/pst.rs:90
        assert sqrt_pst.Polynomial.verify(challenge(), vk, opened.u, r, (v + 1) % curve.R_ORDER, opened.pst_proof,
                                          opened.mipp, t_gt) is False
        key_size = 8 + 8 + sum(8 + (1 << (m_row - k)) * 48 for k in range(m_row)) \
            + 8 + sum(8 + (1 << (m_row - k)) * 96 for k in range(m_row)) + 48 + 96
        rows.append({"power": s, "commit_time": round(commit_ms, 3), "opening_time": round(open_ms, 3),
                     "verification_time": round(verify_ms, 3), "proof_size": serialize.proof_size(opened.pst_proof, opened.mipp),
                     "commiter_key_size": key_size})
        print(rows[-1], flush=True)
        ck.close()
    with open(os.path.join(ROOT, "sqrt_pst.csv"), "w", newline="") as f:
        w = csv.DictWriter(f, fieldnames=list(rows[0].keys()))
        w.writeheader()
        w.writerows(rows)


if __name__ == "__main__":
    main()
