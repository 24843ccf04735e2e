"""Generates tests/golden/pairing_golden.json from oracle/pairing.py (run from the repo root:
`python tests/golden/make_golden_pairing.py`). GT values are stored as 12 hex Fq coefficients in the FLAT order of
the oracle (coefficients of w^0..w^5, each (c0, c1)); points are derived from seeds via the oracle's rand_points."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import bls12_377 as o  # noqa: E402
from oracle import bls12_377_g2 as o2  # noqa: E402
from oracle import pairing as pr  # noqa: E402


def gt_hex(a):
    return [hex(c) for pair in a for c in pair]


def main():
    e = pr.pairing(o.G, o2.G2)
    assert e == pr.pairing_definition(o.G, o2.G2)
    out = {"e_G1_G2": gt_hex(e), "seeded": []}
    for n, s1, s2 in [(1, 101, 201), (2, 102, 202), (3, 103, 203), (7, 104, 204), (8, 105, 205), (9, 106, 206), (33, 107, 207), (64, 108, 208)]:
        ps, dl1 = o.rand_points(n, s1)
        qs, dl2 = o2.rand_points(n, s2)
        res = pr.multi_pairing(ps, qs)
        # closed form: e(G1, G2)^(sum a_i b_i)
        assert res == pr.f12_pow(e, sum(a * b for a, b in zip(dl1, dl2)) % o.R_ORDER)
        out["seeded"].append({"n": n, "g1_seed": s1, "g2_seed": s2, "result": gt_hex(res)})
    with open(os.path.join(ROOT, "tests", "golden", "pairing_golden.json"), "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
