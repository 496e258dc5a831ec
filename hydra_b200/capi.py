"""ctypes binding of the C ABI in include/hydra_b200.h (libhydra_b200.so).

This is the only way Python reaches the CUDA path; there is no CPU fallback. If the
shared library is missing, `load()` raises with the build command.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libhydra_b200.so")
_LIB = None

HB_OK = 0
REPR_SPARSE, REPR_BED, REPR_MIXED = 0, 1, 2
NCCL_ID_BYTES = 128

u32p = C.POINTER(C.c_uint32)
u64p = C.POINTER(C.c_uint64)
i32p = C.POINTER(C.c_int32)
f64p = C.POINTER(C.c_double)
u8p = C.POINTER(C.c_uint8)


class HbConfig(C.Structure):
    _fields_ = [
        ("device", C.c_int32), ("n_ind_raw", C.c_uint32), ("n_na", C.c_uint32), ("na_inds", C.c_void_p),
        ("m_total", C.c_uint32), ("n_tasks_total", C.c_uint32), ("task_first", C.c_uint32), ("n_tasks_local", C.c_uint32),
        ("block_starts", C.c_void_p), ("block_lens", C.c_void_p), ("sync_rate", C.c_uint32), ("n_groups", C.c_uint32),
        ("n_mix", C.c_uint32), ("repr_mode", C.c_int32), ("threshold_fnz", C.c_double), ("n_slices", C.c_uint32),
        ("max_ctas", C.c_uint32), ("model", C.c_uint32), ("reserved", C.c_uint32 * 7),
    ]


class HbBrrTape(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("zmu", "perm", "u", "z", "sigmaG", "pi", "sigmaE", "xI", "zcov", "gnu", "glam", "fh_hyper")]


class HbFhConfig(C.Structure):
    _fields_ = [(n, C.c_double) for n in ("v0L", "v0t", "v0c", "s02c", "tau0")]


class HbBrrIterOut(C.Structure):
    _fields_ = [(n, C.c_double) for n in ("sigmaE", "e_sqn", "epssum", "loop_ms", "iter_ms")] + [
        (n, C.c_uint64) for n in ("n_sync", "n_windows", "n_launches", "nnz_processed", "nnz_updated", "bed_markers", "markers_changed")
    ] + [("phase_cycles", C.c_uint64 * 8), ("windows_ahead", C.c_uint64), ("draws_repeated", C.c_uint64)]


class HbBwTape(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("perm", "p", "sigmaG", "pi", "xI")]


class HbBwIterOut(C.Structure):
    _fields_ = [(n, C.c_double) for n in ("mu", "alpha", "loop_ms", "iter_ms")] + [
        (n, C.c_uint64) for n in ("n_sync", "n_windows", "n_launches", "markers_changed", "density_evals")]


class HydraError(RuntimeError):
    pass


def build(force: bool = False) -> str:
    """Compile libhydra_b200.so for sm_100a (nvcc cross-compiles without a GPU)."""
    args = ["make", "-C", os.path.join(_HERE, "csrc")] + (["-B"] if force else [])
    r = subprocess.run(args, capture_output=True, text=True)
    if r.returncode != 0:
        raise HydraError("building libhydra_b200.so failed:\n" + r.stdout + r.stderr)
    return LIB_PATH


def load() -> C.CDLL:
    global _LIB
    if _LIB is not None:
        return _LIB
    if not os.path.exists(LIB_PATH):
        raise HydraError(f"{LIB_PATH} is missing: run `make -C hydra_b200/csrc` (or __graft_entry__.build()); "
                         "hydra_b200 has no CPU fallback")
    L = C.CDLL(LIB_PATH)
    L.hb_last_error.restype = C.c_char_p
    L.hb_genotype_bytes.restype = C.c_uint64
    L.hb_genotype_bytes.argtypes = [C.c_void_p]
    L.hb_destroy.restype = None
    L.hb_destroy.argtypes = [C.c_void_p]
    _LIB = L
    return L


# every symbol include/hydra_b200.h declares (checked by tests/test_abi.py against the header text)
EXPORTS = [
    "hb_abi_version", "hb_sizeof_config", "hb_sizeof_iter_out", "hb_sizeof_brr_tape", "hb_last_error", "hb_create", "hb_destroy", "hb_get_layout", "hb_get_task_blocks", "hb_stage_bed",
    "hb_stage_sparse", "hb_stage_synth", "hb_stage_finalize", "hb_marker_counts", "hb_marker_stats", "hb_marker_is_bed",
    "hb_genotype_bytes", "hb_export_sparse", "hb_export_bed", "hb_set_epsilon", "hb_get_epsilon", "hb_dot_markers",
    "hb_scaadd_markers", "hb_brr_init", "hb_brr_iteration", "hb_brr_get_hyper", "hb_brr_get_state", "hb_brr_get_state_async", "hb_brr_state_wait", "hb_brr_set_state", "hb_brr_restore_outputs", "hb_brr_save_state", "hb_brr_load_state",
    "hb_brr_get_task_epsilon", "hb_brr_get_task_perm", "hb_brr_get_task_rng", "hb_brr_set_covariates", "hb_brr_get_gamma", "hb_brr_set_group_priors", "hb_brr_set_fh", "hb_brr_get_fh", "hb_comm_get_unique_id", "hb_comm_init", "hb_comm_check_equal",
    "hb_bw_init", "hb_bw_iteration", "hb_bw_set_covariates", "hb_bw_get_gamma", "hb_bw_get_hyper", "hb_bw_marker_stats", "hb_bw_vi_sums", "hb_bw_sum_exp",
    "hb_bw_marginal_likelihoods", "hb_bw_arms_beta",
]


def check(rc: int) -> None:
    if rc != HB_OK:
        raise HydraError(f"hydra_b200 error {rc}: {load().hb_last_error().decode(errors='replace')}")


def ptr(a):
    return None if a is None else C.c_void_p(a.ctypes.data)


def arr(a, dt):
    return None if a is None else np.ascontiguousarray(a, dtype=dt)
