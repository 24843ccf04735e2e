"""ctypes binding of include/testudo_b200.h. There is no fallback: if the CUDA library is missing or no GPU is
present, loading / initialising raises."""
from __future__ import annotations

import atexit
import ctypes
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libtestudo_b200.so")

SCALARS_MONT = 1

_lib = None
_lock = threading.Lock()
_inited = False

c_void_p = ctypes.c_void_p
c_size_t = ctypes.c_size_t
c_ssize_t = ctypes.c_ssize_t
c_uint = ctypes.c_uint
c_int = ctypes.c_int

# every symbol include/testudo_b200.h declares: (restype, argtypes)
SIGNATURES = {
    "tb200_init": (c_int, [c_int]),
    "tb200_init_devices": (c_int, [ctypes.POINTER(c_int), c_int]),
    "tb200_device_count": (c_int, []),
    "tb200_shutdown": (None, []),
    "tb200_last_error": (ctypes.c_char_p, []),
    "tb200_launch_count": (ctypes.c_uint64, []),
    "tb200_reset_launch_count": (None, []),
    "tb200_msm_g1": (c_int, [c_void_p, c_void_p, c_size_t, c_uint, c_void_p]),
    "tb200_msm_g1_dev": (c_int, [c_void_p, c_void_p, c_size_t, c_uint, c_void_p, c_void_p]),
    "tb200_msm_g1_sharded_dev": (c_int, [c_void_p, c_void_p, c_void_p, c_uint, c_void_p]),
    "tb200_srs_load_blinded": (c_int, [c_void_p, c_size_t, c_void_p, c_int, ctypes.POINTER(c_void_p)]),
    "tb200_msm_g1_batch_blinded": (c_int, [c_void_p, c_void_p, c_size_t, c_size_t, c_void_p, c_uint, c_void_p]),
    "tb200_sqrt_pst_commit": (c_int, [c_void_p, c_void_p, c_size_t, c_size_t, c_uint, c_void_p, c_void_p, c_void_p]),
    "tb200_sqrt_pst_commit_strided": (c_int, [c_void_p, c_void_p, c_size_t, c_size_t, c_ssize_t, c_ssize_t, c_uint,
                                              c_void_p, c_void_p, c_void_p]),
    "tb200_host_alloc": (c_int, [c_size_t, ctypes.POINTER(c_void_p)]),
    "tb200_host_alloc_near": (c_int, [c_size_t, c_int, ctypes.POINTER(c_void_p)]),
    "tb200_host_alloc_sharded": (c_int, [c_size_t, c_size_t, ctypes.POINTER(c_void_p)]),
    "tb200_device_numa_node": (c_int, [c_int]),
    "tb200_host_free": (c_int, [c_void_p]),
    "tb200_host_register": (c_int, [c_void_p, c_size_t]),
    "tb200_host_unregister": (c_int, [c_void_p]),
    "tb200_srs_load": (c_int, [c_void_p, c_size_t, c_int, ctypes.POINTER(c_void_p)]),
    "tb200_srs_free": (c_int, [c_void_p]),
    "tb200_srs_size": (c_size_t, [c_void_p]),
    "tb200_msm_g1_batch": (c_int, [c_void_p, c_void_p, c_size_t, c_size_t, c_ssize_t, c_ssize_t, c_uint, c_void_p]),
    "tb200_msm_g1_batch_ptrs": (c_int, [c_void_p, c_void_p, c_size_t, c_size_t, c_uint, c_void_p]),
    "tb200_msm_g1_batch_dev": (c_int, [c_void_p, c_void_p, c_size_t, c_size_t, c_ssize_t, c_ssize_t, c_uint,
                                       c_void_p, c_void_p]),
    "tb200_mipp_g1_begin": (c_int, [c_void_p, c_void_p, c_size_t, c_uint, ctypes.POINTER(c_void_p)]),
    "tb200_mipp_g1_len": (c_size_t, [c_void_p]),
    "tb200_mipp_g1_cross": (c_int, [c_void_p, c_void_p, c_void_p]),
    "tb200_mipp_g1_fold": (c_int, [c_void_p, c_void_p, c_void_p]),
    "tb200_mipp_g1_read": (c_int, [c_void_p, c_void_p, c_void_p]),
    "tb200_mipp_g1_end": (c_int, [c_void_p]),
    "tb200_compress_g1": (c_int, [c_void_p, c_size_t, c_void_p, c_uint]),
    "tb200_msm_g2": (c_int, [c_void_p, c_void_p, c_size_t, c_uint, c_void_p]),
    "tb200_msm_g2_dev": (c_int, [c_void_p, c_void_p, c_size_t, c_uint, c_void_p, c_void_p]),
    "tb200_compress_g2": (c_int, [c_void_p, c_size_t, c_void_p, c_uint]),
    "tb200_mipp_g2_begin": (c_int, [c_void_p, c_size_t, c_uint, c_void_p]),
    "tb200_mipp_g2_len": (c_size_t, [c_void_p]),
    "tb200_mipp_g2_fold": (c_int, [c_void_p, c_void_p]),
    "tb200_mipp_g2_read": (c_int, [c_void_p, c_void_p]),
    "tb200_mipp_g2_end": (c_int, [c_void_p]),
    "tb200_multi_pairing": (c_int, [c_void_p, c_void_p, c_size_t, c_void_p]),
    "tb200_multi_pairing_dev": (c_int, [c_void_p, c_void_p, c_size_t, c_void_p, c_void_p]),
    "tb200_miller_product": (c_int, [c_void_p, c_void_p, c_size_t, c_void_p]),
    "tb200_miller_product_dev": (c_int, [c_void_p, c_void_p, c_size_t, c_void_p, c_void_p]),
    "tb200_gt_product_final_exp": (c_int, [c_void_p, c_size_t, c_void_p]),
    "tb200_gt_product_final_exp_dev": (c_int, [c_void_p, c_size_t, c_void_p, c_void_p]),
    "tb200_mipp_pairing_cross": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p]),
    "tb200_set_pairing_coop_max": (None, [c_int]),
    "tb200_set_pairing_team": (None, [c_int]),
    "tb200_mipp_cross_all": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "tb200_gt_pow": (c_int, [c_void_p, c_void_p, c_size_t, c_uint, c_void_p]),
    "tb200_gt_multi_pow": (c_int, [c_void_p, c_void_p, c_size_t, c_uint, c_void_p]),
    "tb200_multi_pairing_batch": (c_int, [c_void_p, c_void_p, c_size_t, c_size_t, c_void_p]),
    "tb200_msm_g1_each": (c_int, [c_void_p, c_void_p, c_size_t, c_size_t, c_uint, c_void_p]),
    "tb200_msm_g1_rows": (c_int, [c_void_p, c_void_p, c_void_p, c_size_t, c_uint, c_void_p]),
    "tb200_msm_g1_begin": (c_int, [c_void_p, c_void_p, c_size_t, c_uint, c_void_p]),
    "tb200_msm_g1_end": (c_int, [c_void_p, c_void_p]),
    "tb200_pst_open_g1": (c_int, [c_void_p, c_size_t, c_void_p, c_void_p, c_uint, c_void_p]),
    "tb200_pst_open_g2": (c_int, [c_void_p, c_size_t, c_void_p, c_void_p, c_uint, c_void_p]),
    "tb200_pst_open_g1_begin": (c_int, [c_void_p, c_size_t, c_void_p, c_void_p, c_uint, ctypes.POINTER(c_void_p)]),
    "tb200_pst_open_g2_begin": (c_int, [c_void_p, c_size_t, c_void_p, c_void_p, c_uint, ctypes.POINTER(c_void_p)]),
    "tb200_pst_open_end": (c_int, [c_void_p, c_void_p]),
    "tb200_dev_alloc": (c_int, [c_size_t, ctypes.POINTER(c_void_p)]),
    "tb200_dev_free": (c_int, [c_void_p]),
    "tb200_dev_upload": (c_int, [c_void_p, c_void_p, c_size_t]),
    "tb200_dev_download": (c_int, [c_void_p, c_void_p, c_size_t]),
    "tb200_stream_sync": (c_int, []),
    "tb200_fr_chis": (c_int, [c_void_p, c_size_t, c_void_p]),
    "tb200_fr_subset_products": (c_int, [c_void_p, c_size_t, c_void_p]),
    "tb200_fr_matvec": (c_int, [c_void_p, c_size_t, c_size_t, c_void_p, c_void_p]),
    "tb200_fr_matvec_dev": (c_int, [c_void_p, c_size_t, c_size_t, c_void_p, c_void_p, c_void_p]),
    "tb200_g1_sum": (c_int, [c_void_p, c_size_t, c_void_p]),
    "tb200_g1_sum_dev": (c_int, [c_void_p, c_size_t, c_void_p, c_void_p]),
    "tb200_g1_outer_sum_dev": (c_int, [c_void_p, c_size_t, c_void_p, c_size_t, c_void_p, c_void_p]),
    "tb200_set_profiling": (None, [c_int]),
    "tb200_stage_ms": (ctypes.c_double, [ctypes.c_char_p]),
    "tb200_last_geometry": (c_int, [ctypes.POINTER(c_int), ctypes.POINTER(c_int), ctypes.POINTER(ctypes.c_uint64),
                                    ctypes.POINTER(ctypes.c_uint64), ctypes.POINTER(c_int)]),
    "tb200_set_window_bits": (None, [c_int]),
    "tb200_set_accumulate_mode": (None, [c_int]),
    "tb200_set_pass_entries_max": (None, [ctypes.c_uint64]),
    "tb200_set_shard_min": (None, [c_size_t]),
    "tb200_set_commit_pipeline": (None, [c_int]),
    "tb200_set_small_msm_max": (None, [c_int]),
    "tb200_set_msm_overlap": (None, [c_int]),
    "tb200_set_host_upload": (c_int, [c_int, ctypes.POINTER(c_int), c_int]),
    "tb200_int_pipe_peak": (c_int, [c_int, c_int, ctypes.POINTER(ctypes.c_double)]),
    "tb200_test_fq_mul": (c_int, [c_void_p, c_void_p, c_size_t, c_void_p]),
    "tb200_test_fq_addsub": (c_int, [c_void_p, c_void_p, c_size_t, c_void_p, c_void_p]),
    "tb200_test_g1_add": (c_int, [c_void_p, c_void_p, c_size_t, c_void_p]),
    "tb200_test_g1_mul": (c_int, [c_void_p, c_void_p, c_size_t, c_void_p]),
    "tb200_test_g2_add": (c_int, [c_void_p, c_void_p, c_size_t, c_void_p]),
    "tb200_test_fq12_op": (c_int, [c_int, c_void_p, c_void_p, c_size_t, c_void_p]),
    "tb200_test_g2_mul": (c_int, [c_void_p, c_void_p, c_size_t, c_void_p]),
    "tb200_poseidon_new": (c_int, [c_int, c_uint, c_uint, ctypes.c_uint64, c_uint, c_uint, c_void_p, c_void_p,
                                   ctypes.POINTER(c_void_p)]),
    "tb200_poseidon_reset": (c_int, [c_void_p]),
    "tb200_poseidon_absorb_bytes": (c_int, [c_void_p, c_void_p, c_size_t]),
    "tb200_poseidon_append_words": (c_int, [c_void_p, c_void_p, c_size_t]),
    "tb200_poseidon_absorb_native": (c_int, [c_void_p, c_void_p, c_size_t]),
    "tb200_poseidon_squeeze_native": (c_int, [c_void_p, c_void_p, c_size_t]),
    "tb200_poseidon_squeeze_fr": (c_int, [c_void_p, c_void_p]),
    "tb200_poseidon_limbs": (c_int, [c_void_p]),
    "tb200_poseidon_free": (c_int, [c_void_p]),
}


class EngineError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"testudo_b200 error {code}: {message}")
        self.code = code


def load() -> ctypes.CDLL:
    """dlopen the library and bind every declared symbol (no CUDA call is made)."""
    global _lib
    with _lock:
        if _lib is None:
            if not os.path.exists(LIB_PATH):
                raise RuntimeError(
                    f"{LIB_PATH} is missing: build it with `python -m testudo_b200.build` "
                    "(testudo_b200 has no CPU fallback)")
            lib = ctypes.CDLL(LIB_PATH)
            for name, (res, args) in SIGNATURES.items():
                fn = getattr(lib, name)  # AttributeError if the .so does not export a declared symbol
                fn.restype = res
                fn.argtypes = args
            _lib = lib
            atexit.register(lib.tb200_shutdown)  # joins the per-GPU worker threads, destroys the NCCL clique
    return _lib


def check(rc: int) -> None:
    if rc != 0:
        raise EngineError(rc, load().tb200_last_error().decode("utf-8", "replace"))


def init(device: int = -1) -> ctypes.CDLL:
    """Create the CUDA context on `device` (raises without a GPU). TB200_DEVICES="0,1,..." in the environment makes the
    first initialisation a multi-device one (tb200_init_devices)."""
    global _inited
    lib = load()
    if not _inited:
        env = os.environ.get("TB200_DEVICES", "").strip()
        if env and device < 0:
            init_devices([int(x) for x in env.split(",") if x.strip()])
        else:
            check(lib.tb200_init(device))
        _inited = True
    return lib


def init_devices(devices) -> ctypes.CDLL:
    """One process, several GPUs: `devices[0]` is the primary (tb200_init_devices)."""
    global _inited
    lib = load()
    arr = (c_int * len(devices))(*devices)
    check(lib.tb200_init_devices(arr, len(devices)))
    _inited = True
    return lib


def device_count() -> int:
    return load().tb200_device_count()


def engine() -> ctypes.CDLL:
    return init()
