"""Test infrastructure only: CPU oracle for the BLS12-377 G1 MSM hot path (see bls12_377.py, cpu_msm.c)."""
