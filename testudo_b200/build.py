"""Builds the engine's CUDA shared library in-tree with nvcc for sm_100a (cross-compiles without a GPU).

The engine is four translation units (csrc/engine.h explains the split); they compile in parallel and are linked into
one libtestudo_b200.so. Only a unit whose sources changed is recompiled."""
from __future__ import annotations

import os
import shutil
import subprocess
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB_DIR = os.path.join(HERE, "lib")
OBJ_DIR = os.path.join(LIB_DIR, "obj")
LIB = os.path.join(LIB_DIR, "libtestudo_b200.so")
API_HEADER = os.path.join("..", "..", "include", "testudo_b200.h")
_FIELD = ["mont.cuh", "g1.cuh", "g1_fast.cuh", "digits.cuh", "kernels.cuh"]
_G2 = _FIELD + ["g2.cuh", "kernels_g2.cuh"]
# unit -> the headers it includes (directly or not)
UNITS = {
    "engine_core.cu": [],
    "engine_g1.cu": _FIELD + ["kernels_smem.cuh", "kernels_small.cuh"],
    "engine_g2.cu": _G2,
    "engine_pairing.cu": _G2 + ["fq12.cuh", "fq12_coop.cuh", "fq12_consts.inc", "kernels_pairing.cuh"],
    "poseidon_host.cpp": [],   # host-only: the Fiat-Shamir sponge (CPU code in the reference too)
}
COMMON = ["engine.h", "glv_host.h", API_HEADER]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC",
]


def nvcc() -> str:
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found: the engine has no non-CUDA build")
    return exe


def _obj(unit: str) -> str:
    return os.path.join(OBJ_DIR, os.path.splitext(unit)[0] + ".o")


def _unit_stale(unit: str) -> bool:
    obj = _obj(unit)
    if not os.path.exists(obj):
        return True
    t = os.path.getmtime(obj)
    deps = [os.path.join(CSRC, f) for f in [unit] + UNITS[unit] + COMMON]
    return any(os.path.exists(d) and os.path.getmtime(d) > t for d in deps)


def is_stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(_unit_stale(u) or os.path.getmtime(_obj(u)) > t for u in UNITS)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not is_stale():
        return LIB
    os.makedirs(OBJ_DIR, exist_ok=True)
    todo = [u for u in UNITS if force or _unit_stale(u)]

    def compile_unit(unit: str) -> None:
        cmd = [nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", unit, "-o", _obj(unit)]
        subprocess.check_call(cmd, cwd=CSRC)

    with ThreadPoolExecutor(max_workers=max(1, len(todo))) as pool:
        list(pool.map(compile_unit, todo))
    # NCCL is resolved with dlopen when a second device is added (engine_core.cu): no link-time dependency
    cmd = [nvcc(), "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB] + [_obj(u) for u in UNITS] + ["-ldl"]
    subprocess.check_call(cmd, cwd=CSRC)
    return LIB


if __name__ == "__main__":
    import sys

    print(build(force="--force" in sys.argv or len(sys.argv) == 1, verbose="-v" in sys.argv))
