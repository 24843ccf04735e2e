"""Generates tests/golden/g2_golden.json with the big-integer G2 oracle (oracle/bls12_377_g2.py).

The reference holds no golden vectors for its G2 MSMs either: seeded inputs -> naive double-and-add
MSM (the mathematical definition), cross-checked against the closed-form discrete-log result.
Run:  python tests/golden/make_golden_g2.py
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import bls12_377 as o  # noqa: E402
from oracle import bls12_377_g2 as o2  # noqa: E402


def hx(p):
    return None if p is None else [hex(c) for c in (p[0][0], p[0][1], p[1][0], p[1][1])]


def main():
    out = {"curve_kats": {
        "G2": hx(o2.G2), "2G2": hx(o2.add(o2.G2, o2.G2)), "3G2": hx(o2.add(o2.add(o2.G2, o2.G2), o2.G2)),
        "(r-1)G2": hx(o2.mul(o2.R_ORDER - 1, o2.G2)), "rG2_is_inf": o2.add(o2.mul(o2.R_ORDER - 1, o2.G2), o2.G2) is None,
        "12345G2": hx(o2.mul(12345, o2.G2)),
    }, "seeded": [], "explicit": [], "edge": []}
    for n in (1, 2, 3, 31, 33, 256, 1024):
        pts, dl = o2.rand_points(n, 7000 + n)
        sc = o.rand_scalars(n, 8000 + n)
        exp = o2.msm_by_dlog(dl, sc)
        if n <= 33:
            assert o2.msm_naive(pts, sc) == exp
        out["seeded"].append({"n": n, "points_seed": 7000 + n, "scalars_seed": 8000 + n, "result": hx(exp)})
        print("seeded", n)
    pts, _ = o2.rand_points(5, 9001)
    sc = o.rand_scalars(5, 9002)
    out["explicit"].append({"points": [hx(p) for p in pts], "scalars": [hex(s) for s in sc],
                            "result": hx(o2.msm_naive(pts, sc))})
    p8, _ = o2.rand_points(8, 9100)
    r = o.R_ORDER
    edge = [
        ("all_zero_scalars", p8, [0] * 8),
        ("all_one_scalars", p8, [1] * 8),
        ("r_minus_one", p8[:3], [r - 1] * 3),
        ("top_bit_252", p8[:4], [(1 << 252) + i for i in range(4)]),
        ("identity_bases", [None, p8[1], None, p8[3]], o.rand_scalars(4, 9101)),
        ("duplicate_bases", [p8[0]] * 6, o.rand_scalars(6, 9102)),
        ("base_and_negation", [p8[2], o2.neg(p8[2])], [12345, 12345]),
        ("same_base_same_bucket", [p8[4], p8[4]], [7, 7]),
    ]
    for name, pts, sc in edge:
        out["edge"].append({"name": name, "points": [hx(p) for p in pts], "scalars": [hex(s) for s in sc],
                            "result": hx(o2.msm_naive(pts, sc))})
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "g2_golden.json"), "w") as f:
        json.dump(out, f, indent=0)
    print("written")


if __name__ == "__main__":
    main()
