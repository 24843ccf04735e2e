"""CPU tests: the oracle's restatement of the reference's sqrt-PST prover (oracle/sqrt_pst.py) and verifier
(oracle/verifier.py) run the reference's own round-trip test `check_sqrt_poly_commit` (src/sqrt_pst.rs:297-342) at sizes
Python finishes in seconds -- this pins the two restatements against each other (odd and even num_vars, as the
reference does) before the GPU path is compared with them (tests/test_gpu_sqrt_pst.py)."""
import hashlib

import pytest

from oracle import bls12_377 as o
from oracle import bls12_377_g2 as o2
from oracle import pairing as pr
from oracle import sqrt_pst as osp
from oracle import verifier as ver


def transcript():
    state = hashlib.sha256(b"oracle-roundtrip")

    def challenge(label, values):
        state.update(label)
        for kind, v in values:
            words = o.affine_to_words(v) if kind == "g1" else (o2.affine_to_words(v) if kind == "g2" else pr.to_words(v))
            state.update(b"".join(int(w).to_bytes(8, "little") for w in words))
        return int.from_bytes(state.digest(), "little") % o.R_ORDER or 1

    return challenge


@pytest.mark.parametrize("nv", [3, 4])
def test_check_sqrt_poly_commit_oracle_roundtrip(nv):
    m_row = nv - nv // 2
    t = o.rand_scalars(m_row, 2900 + nv)
    ck = osp.setup_ck(t)
    vk = ver.setup_vk(t)
    z = o.rand_scalars(1 << nv, 2910 + nv)
    r = o.rand_scalars(nv, 2920 + nv)
    poly = osp.Polynomial(z)
    v = poly.eval(r)
    # check_sqrt_poly_eval (src/sqrt_pst.rs:277-295): the sqrt layout evaluates like the dense polynomial
    assert v == sum(zi * osp.get_chi_i(r, i) for i, zi in enumerate(z)) % o.R_ORDER
    comm_list, T = poly.commit(ck)
    U, pst_proof, mipp_proof = poly.open(transcript(), comm_list, ck, r, T)
    assert ver.sqrt_pst_verify(vk, transcript(), U, r, v, pst_proof, mipp_proof, T) is True
    assert ver.sqrt_pst_verify(vk, transcript(), U, r, (v + 1) % o.R_ORDER, pst_proof, mipp_proof, T) is False
    bad = dict(mipp_proof)
    bad["final_a"] = o.add(bad["final_a"], o.G)
    assert ver.sqrt_pst_verify(vk, transcript(), U, r, v, pst_proof, bad, T) is False
