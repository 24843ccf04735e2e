#!/usr/bin/env python
"""Mirror of the reference's only live benchmark, benches/pst.rs (BASELINE configs[0]), for the G1 stages this engine
covers. Same loop (`for s in [4, 5, 20, 27]`, benches/pst.rs:26 -- 27 is replaced by 26, the BASELINE size) and the
same CSV columns (benches/pst.rs:13-21,93-96): power, commit_time, opening_time, verification_time, proof_size,
commiter_key_size -- times in ms.

  commit_time   = `Polynomial::commit` row stage `comm_list` (src/sqrt_pst.rs:121-125) through the host-facing batched
                  call (host scalars, H2D inside). The pairing product `ipp` (src/sqrt_pst.rs:131-144) is out of scope.
  opening_time  = G1 work of `Polynomial::open` (src/sqrt_pst.rs:168-230): get_q on the device, M2 `msm_unchecked`,
                  M3 `commit(q)`, and the device-resident MIPP G1 loop; G2 openings / pairings are out of scope.
  verification_time, proof_size = n/a (verifier and G2 side out of scope) -> empty.
  commiter_key_size = bytes of ck.powers_of_g[0] in ark's uncompressed encoding (96 B per point).
Synthetic inputs like the reference (`F::rand(test_rng)`, MultilinearPC::setup): uniform scalars and an SRS of
subgroup points with known discrete logs, generated on the GPU.
"""
import csv
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from testudo_b200 import _lib, curve, sqrt_pst  # noqa: E402
from testudo_b200.synthetic import make_bases_dev, make_scalars_dev  # noqa: E402


def main():
    sizes = [int(a) for a in sys.argv[1:]] or [4, 5, 20, 26]
    _lib.init()
    rows = []
    state = {"k": 0x1234567}

    def challenge(label, pts):  # stand-in for the Poseidon transcript (out of scope)
        state["k"] = (state["k"] * 6364136223846793005 + 1442695040888963407) % curve.R_ORDER
        return state["k"] | 1

    for s in sizes:
        m_row = s - s // 2
        z = make_scalars_dev(1 << s, seed=s).cpu().numpy().view(np.uint64)     # Montgomery-form Fr, like ark memory
        srs = make_bases_dev(1 << m_row, seed=100 + s).cpu().numpy().view(np.uint64)
        ck = sqrt_pst.CommitterKey.from_points(srs)                              # setup + trim
        pl = sqrt_pst.Polynomial.from_evaluations(z)
        r = [int.from_bytes(np.random.default_rng(s + i).bytes(31), "little") % curve.R_ORDER for i in range(s)]
        pl.commit(ck)                                                            # warm-up (tables, arena)
        t0 = time.perf_counter()
        comm_list, t = pl.commit(ck)
        commit_ms = (time.perf_counter() - t0) * 1e3
        t0 = time.perf_counter()
        pl.open(challenge, comm_list, ck, r)
        open_ms = (time.perf_counter() - t0) * 1e3
        rows.append({"power": s, "commit_time": round(commit_ms, 3), "opening_time": round(open_ms, 3),
                     "verification_time": "", "proof_size": "", "commiter_key_size": 96 * (1 << m_row)})
        print(rows[-1], flush=True)
        ck.close()
    with open(os.path.join(ROOT, "sqrt_pst.csv"), "w", newline="") as f:
        w = csv.DictWriter(f, fieldnames=list(rows[0].keys()))
        w.writeheader()
        w.writerows(rows)


if __name__ == "__main__":
    main()
