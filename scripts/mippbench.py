"""Timing of the device-resident MIPP G1 loop (src/mipp.rs:58-120) at the C3 shape: 2^13 commitments."""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from testudo_b200 import _lib, mipp, curve
from testudo_b200.synthetic import make_bases_dev, make_scalars_dev
lib = _lib.init()
logn = int(sys.argv[1]) if len(sys.argv) > 1 else 13
n = 1 << logn
a = make_bases_dev(n, seed=5).cpu().numpy().view(np.uint64)
y = make_scalars_dev(n, seed=6).cpu().numpy().view(np.uint64)   # treated as Montgomery-form Fr
st = {"k": 12345}
def challenge(label, pts):
    st["k"] = (st["k"] * 6364136223846793005 + 1442695040888963407) % curve.R_ORDER
    return st["k"] | 1
for rep in range(2):
    t0 = time.time()
    proof = mipp.MippProofG1.prove(challenge, a, y, a[0])
    dt = time.time() - t0
print(json.dumps({"mipp_g1_prove_n": n, "rounds": len(proof.comms_u), "ms": round(dt * 1e3, 2),
                  "ms_per_round": round(dt * 1e3 / len(proof.comms_u), 2)}))
