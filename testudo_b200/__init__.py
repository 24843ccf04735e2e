"""testudo_b200 -- B200-native BLS12-377 G1 MSM engine behind Testudo's commitment call surface.

Host-side mirrors of the reference interfaces (all arithmetic runs in the CUDA library, see csrc/):
  msm.py          ark-ec `VariableBaseMSM::{msm, msm_unchecked, msm_bigint}`
  sqrt_pst.py     `sqrt_pst::Polynomial::{from_evaluations, commit, open}` (G1 work)
  mipp.py         `mipp::{multiexponentiation, compress, MippProof::prove}` (G1 work)
  commitments.py  `commitments::{MultiCommitGens, PedersenCommit}`
  parallel.py     one-process-per-GPU sharding + NCCL all-gather of partial results
"""
from ._lib import EngineError, SCALARS_MONT, init, load  # noqa: F401

__all__ = ["EngineError", "SCALARS_MONT", "init", "load"]
