"""ctypes binding of libb200ref.so (include/b200ref.h).

This is the reference-side stub a maintainer would add: plain
pointers and sizes, device memory owned by the caller (torch tensors here).
"""

from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

_LIB_PATH = Path(__file__).resolve().parent / "_lib" / "libb200ref.so"

ERR_ARG, ERR_WORKSPACE, ERR_CUDA, ERR_NONFINITE = -1, -2, -3, -4
SUMMARY_K = 20
S_FLAGS = 15  # B200REF_S_FLAGS: bit0 mean passed, bit1 std passed
MAX_QUANTILES = 8
ESS_MODES = {"first_negative": 0, "geyer": 1}
GLM_FAMILIES = {"bernoulli_logit": 0, "binomial_logit": 1, "poisson_log": 2, "normal": 3}

EXPORTS = (
    "b200ref_version", "b200ref_last_error", "b200ref_kernel_launches",
    "b200ref_profile", "b200ref_profile_read", "b200ref_profile_name",
    "b200ref_workspace_bytes", "b200ref_workspace_min_bytes",
    "b200ref_ranks_f64", "b200ref_diagnostics_f64", "b200ref_diagnostics_ex_f64", "b200ref_stats_f64",
    "b200ref_compare_f64", "b200ref_checks_f64",
    "b200ref_full_pass_workspace_bytes", "b200ref_full_pass_f64", "b200ref_full_pass_host_f64", "b200ref_synth_fill_f64",
    "b200ref_glm_workspace_bytes", "b200ref_glm_logistic_lp_f64", "b200ref_glm_lp_f64", "b200ref_glm_lp_ex_f64",
    "b200ref_row_order_i64", "b200ref_gather_rows_f64",
    "b200ref_rows_workspace_bytes", "b200ref_diagnostics_layout_f64", "b200ref_transpose_rows_f64",
    "b200ref_chain_stats_f64",
)
GLM_REUSE_DESIGN = 1
LAYOUTS = {"pcn": 0, "rows": 1}


class Outputs(C.Structure):
    """struct b200ref_outputs"""

    _fields_ = [(n, C.c_void_p) for n in (
        "rhat", "rhat_bulk", "rhat_tail", "ess_bulk", "ess_tail", "lag_bulk", "lag_tail",
        "mean", "std", "quantiles", "order_stats", "median")]


class NativeError(RuntimeError):
    pass


_lib = None


def lib_path() -> Path:
    return Path(os.environ.get("B200REF_LIB", str(_LIB_PATH)))


def load() -> C.CDLL:
    """Load the shared library; fail loudly if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    path = lib_path()
    if not path.exists():
        raise NativeError(
            f"{path} not found: the CUDA extension is not built (run `python -c 'import __graft_entry__ as g; "
            "g.build()'`).  There is no CPU fallback."
        )
    lib = C.CDLL(str(path))
    vp, i64, sz, dbl, u64 = C.c_void_p, C.c_int64, C.c_size_t, C.c_double, C.c_uint64
    lib.b200ref_version.restype = C.c_int
    lib.b200ref_last_error.restype = C.c_char_p
    lib.b200ref_kernel_launches.restype = u64
    lib.b200ref_profile.argtypes = [C.c_int]
    lib.b200ref_profile_read.argtypes = [C.POINTER(dbl), C.POINTER(u64), C.c_int, C.c_int]
    lib.b200ref_profile_name.restype = C.c_char_p
    lib.b200ref_profile_name.argtypes = [C.c_int]
    lib.b200ref_workspace_bytes.restype = sz
    lib.b200ref_workspace_bytes.argtypes = [i64, i64, i64]
    lib.b200ref_workspace_min_bytes.restype = sz
    lib.b200ref_workspace_min_bytes.argtypes = [i64, i64]
    lib.b200ref_ranks_f64.argtypes = [vp, i64, i64, i64, vp, vp, vp, vp, vp, sz, vp]
    lib.b200ref_diagnostics_f64.argtypes = [vp, i64, i64, i64, C.POINTER(dbl), C.c_int, C.POINTER(Outputs), vp, sz, vp]
    lib.b200ref_diagnostics_ex_f64.argtypes = [vp, i64, i64, i64, C.POINTER(dbl), C.c_int, C.c_int, C.POINTER(Outputs), vp, sz, vp]
    lib.b200ref_stats_f64.argtypes = [vp, i64, i64, C.POINTER(dbl), C.c_int, C.POINTER(Outputs), vp, sz, vp]
    lib.b200ref_compare_f64.argtypes = [vp, vp, i64, dbl, vp, vp, vp, vp]
    lib.b200ref_checks_f64.argtypes = [vp, vp, i64, C.POINTER(C.c_uint8), vp]
    lib.b200ref_full_pass_workspace_bytes.restype = sz
    lib.b200ref_full_pass_workspace_bytes.argtypes = [i64, i64, i64, i64]
    lib.b200ref_full_pass_f64.argtypes = [vp, i64, i64, i64, vp, i64, dbl, dbl, vp, C.POINTER(C.c_uint8), vp, sz, vp]
    lib.b200ref_full_pass_host_f64.argtypes = [vp, i64, i64, i64, vp, i64, dbl, dbl, vp, C.POINTER(C.c_uint8), vp, sz]
    lib.b200ref_synth_fill_f64.argtypes = [vp, i64, i64, i64, i64, u64, C.c_int, vp]
    lib.b200ref_glm_workspace_bytes.restype = sz
    lib.b200ref_glm_workspace_bytes.argtypes = [i64, i64, i64]
    lib.b200ref_glm_logistic_lp_f64.argtypes = [vp, vp, vp, i64, i64, i64, vp, vp, sz, vp]
    lib.b200ref_glm_lp_f64.argtypes = [C.c_int, vp, vp, vp, vp, vp, i64, i64, i64, vp, vp, sz, vp]
    lib.b200ref_glm_lp_ex_f64.argtypes = [C.c_int, vp, vp, vp, vp, vp, i64, i64, i64, C.c_int, vp, vp, sz, vp]
    lib.b200ref_row_order_i64.argtypes = [vp, vp, i64, C.POINTER(C.c_int), vp]
    lib.b200ref_gather_rows_f64.argtypes = [vp, vp, i64, i64, vp, vp]
    lib.b200ref_rows_workspace_bytes.restype = sz
    lib.b200ref_rows_workspace_bytes.argtypes = [i64, i64, i64]
    lib.b200ref_diagnostics_layout_f64.argtypes = [vp, C.c_int, i64, i64, i64, C.POINTER(dbl), C.c_int, C.c_int,
                                                   C.POINTER(Outputs), vp, sz, vp]
    lib.b200ref_chain_stats_f64.argtypes = [vp, vp, i64, i64, i64, C.c_int, C.POINTER(Outputs), vp, sz, vp]
    lib.b200ref_transpose_rows_f64.argtypes = [vp, i64, i64, i64, i64, vp, vp]
    for name in EXPORTS:  # every symbol include/b200ref.h declares must be exported
        getattr(lib, name)
    _lib = lib
    return lib


def check(rc: int) -> None:
    """Translate a negative return code into the exception the reference path would raise."""
    if rc == 0:
        return
    msg = load().b200ref_last_error().decode()
    if rc == ERR_NONFINITE:
        raise ValueError(msg)
    if rc == ERR_ARG:
        raise ValueError(msg)
    if rc == ERR_WORKSPACE:
        raise MemoryError(msg)
    raise NativeError(msg)


def kernel_launches() -> int:
    return int(load().b200ref_kernel_launches())


def profile(enable: bool) -> None:
    load().b200ref_profile(1 if enable else 0)


def profile_read(reset: bool = True) -> dict[str, tuple[float, int]]:
    """{kernel name: (total ms, launches)} accumulated while profiling was enabled."""
    lib = load()
    n = lib.b200ref_profile_read(None, None, 0, 0)
    ms = (C.c_double * n)()
    cnt = (C.c_uint64 * n)()
    lib.b200ref_profile_read(ms, cnt, n, 1 if reset else 0)
    return {lib.b200ref_profile_name(i).decode(): (float(ms[i]), int(cnt[i])) for i in range(n)}
