"""Generates tests/golden/msm_golden.json with the big-integer oracle (oracle/bls12_377.py).

The reference holds no golden vectors for its MSM path, so these are created here:
seeded inputs -> naive double-and-add MSM (the mathematical definition), cross-checked against the
closed-form discrete-log oracle and against the pure-Python restatement of ark's Pippenger.
Run:  python tests/golden/make_golden.py
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import bls12_377 as o  # noqa: E402
import helpers as h  # noqa: E402


def main():
    out = {"curve_kats": {}, "seeded": [], "explicit": [], "edge": []}
    out["curve_kats"] = {
        "G": h.pt_hex(o.G),
        "2G": h.pt_hex(o.add(o.G, o.G)),
        "3G": h.pt_hex(o.add(o.add(o.G, o.G), o.G)),
        "(r-1)G": h.pt_hex(o.mul(o.R_ORDER - 1, o.G)),
        "rG_is_inf": o.mul(o.R_ORDER, o.G) is None,
        "12345G": h.pt_hex(o.mul(12345, o.G)),
    }
    for n in (1, 2, 3, 31, 32, 33, 256, 1024, 4096):
        pts, dl = o.rand_points(n, 1000 + n)
        sc = o.rand_scalars(n, 2000 + n)
        exp = o.msm_by_dlog(dl, sc)
        if n <= 1024:
            assert o.msm_naive(pts, sc) == exp
        if n <= 256:
            assert o.msm_pippenger(pts, sc) == exp
        out["seeded"].append({"n": n, "points_seed": 1000 + n, "scalars_seed": 2000 + n, "result": h.pt_hex(exp)})
        print("seeded", n)
    # fully explicit small vectors (guards against generator drift)
    for n in (1, 5, 33):
        pts, _ = o.rand_points(n, 3000 + n)
        sc = o.rand_scalars(n, 4000 + n)
        out["explicit"].append({
            "points": [h.pt_hex(p) for p in pts],
            "scalars": [hex(s) for s in sc],
            "result": h.pt_hex(o.msm_naive(pts, sc)),
        })
    for name, pts, sc in h.edge_case_inputs():
        exp = o.msm_naive(pts, sc)
        assert o.msm_pippenger(pts, sc) == exp, name
        out["edge"].append({"name": name, "points": [h.pt_hex(p) for p in pts], "scalars": [hex(s) for s in sc],
                            "result": h.pt_hex(exp)})
    # sqrt_pst layout: Z (2^n scalars) viewed as 2^m_row x 2^m_col, row commitments over a shared SRS
    for nv in (4, 5, 6):
        m_col = nv // 2
        m_row = nv - m_col
        srs, _ = o.rand_points(1 << m_row, 5000 + nv)
        z = o.rand_scalars(1 << nv, 6000 + nv)
        rows = []
        for i in range(1 << m_col):
            col = [z[(j << m_col) | i] for j in range(1 << m_row)]  # src/sqrt_pst.rs:58
            rows.append(h.pt_hex(o.msm_naive(srs, col)))
        out.setdefault("sqrt_rows", []).append({"num_vars": nv, "srs_seed": 5000 + nv, "z_seed": 6000 + nv, "rows": rows})
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "msm_golden.json"), "w") as f:
        json.dump(out, f, indent=1)
    print("written")


if __name__ == "__main__":
    main()
