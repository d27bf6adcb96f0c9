"""ctypes binding of ``libplbert_b200.so`` (C ABI declared in ``include/plbert_b200.h``).

The product path has NO fallback: if the shared library is missing or a call fails, a ``PlbError`` is raised.
Pointers cross the boundary as plain integers (``tensor.data_ptr()``), the stream as the raw ``cudaStream_t``.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import c_char_p, c_double, c_float, c_int, c_long, c_longlong, c_uint, c_ulonglong, c_void_p

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
# PLBERT_B200_LIB points at an alternative build of the same ABI (profiles/ uses it for instrumented / experimental variants)
LIB_PATH = os.environ.get("PLBERT_B200_LIB") or os.path.join(PKG_DIR, "libplbert_b200.so")

GEMM_FWD, GEMM_DGRAD, GEMM_WGRAD = 0, 1, 2
EPI_NONE, EPI_RESIDUAL, EPI_GELU, EPI_GELU_BWD, EPI_ROWDOT = 0, 1, 2, 3, 4
ABI_VERSION = 1


class PlbError(RuntimeError):
    pass


P = c_void_p
_GEMM_ARGS = [c_int, c_int, c_int, c_int, c_int, P, c_int, P, c_int, P, c_int, P, P, c_int, P, c_int, P]

# name -> argtypes; must list every function declared in include/plbert_b200.h (tests/test_abi.py checks this)
SIGNATURES = {
    "plb_abi_version": [],
    "plb_last_error_string": [],
    "plb_device_sm_count": [],
    "plb_set_launch_overlap": [c_int],
    "plb_fill_zero": [P, ctypes.c_ulonglong, P],
    "plb_length_to_mask": [P, c_int, c_int, P, P],
    "plb_gemm_bf16": _GEMM_ARGS,
    "plb_gemm_bf16_check": _GEMM_ARGS,
    "plb_gemm_wgrad_allreduce": [c_int, c_int, c_int, P, c_int, P, c_int, P, c_int, P],
    "plb_embed_ln_fwd": [P, P, c_int, c_int, P, P, P, P, P, c_float, P, P, P],
    "plb_embed_ln_bwd": [P, P, P, c_int, c_int, c_int, P, P, P, P, P, c_int, P, P, P, P, P, P],
    "plb_ln_fwd": [P, c_int, c_int, P, P, c_float, P, P, P],
    "plb_ln_bwd": [P, P, P, P, c_int, c_int, P, P, P, P, P],
    "plb_attn_fwd_tc": [P, c_int, P, P, P, c_int, c_int, c_int, c_int, c_int, P, c_int, P, P],
    "plb_attn_bwd_tc": [P, c_int, P, P, P, c_int, c_int, c_int, c_int, c_int, P, P, c_int, P, P, P, c_int, P, P],
    "plb_attn_bwd_fused_workspace_bytes": [c_int, c_int],
    "plb_attn_bwd_fused": [P, c_int, P, P, P, c_int, c_int, c_int, c_int, c_int, P, P, c_int, P, P, P, c_int, P, P, c_longlong, P],
    "plb_attn_fwd_tc_dropout": [P, c_int, P, P, P, c_int, c_int, c_int, c_int, c_int, P, c_int, P, c_float, c_ulonglong, c_uint, P],
    "plb_attn_bwd_fused_dropout": [P, c_int, P, P, P, c_int, c_int, c_int, c_int, c_int, P, P, c_int, P, P, P, c_int, P, P, c_longlong,
                                   c_float, c_ulonglong, c_uint, P],
    "plb_dropout": [P, c_int, P, c_int, P, c_int, c_int, c_int, c_float, c_ulonglong, c_uint, P],
    "plb_gather_rows": [P, c_int, P, c_int, c_int, P, c_int, P],
    "plb_scatter_add_rows": [P, c_int, P, c_int, c_int, P, c_int, P],
    "plb_masked_ce": [P, c_int, c_int, c_int, P, P, P, P, c_int, P],
    "plb_colsum": [P, c_int, c_int, c_int, P, P],
    "plb_adamw_step": [P, P, P, P, P, c_long, c_int, c_double, c_double, c_double, c_double, c_double, c_double, c_int, P],
    "plb_cast_bf16": [P, P, c_long, P],
}

_lib = None


def load() -> ctypes.CDLL:
    """Load the CUDA library (once).  Raises PlbError when it has not been built — there is no CPU fallback."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise PlbError(
            f"{LIB_PATH} not found: the sm_100a CUDA library has not been built. "
            "Run `python -m plbert_b200.build` (needs nvcc); plbert_b200 has no CPU or PyTorch fallback.")
    lib = ctypes.CDLL(LIB_PATH)
    for name, argtypes in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError here means header / library / binding drifted apart
        fn.argtypes = argtypes
        fn.restype = c_char_p if name == "plb_last_error_string" else c_int
    if lib.plb_abi_version() != ABI_VERSION:
        raise PlbError(f"ABI mismatch: library {lib.plb_abi_version()} vs binding {ABI_VERSION}; rebuild the library")
    _lib = lib
    return lib


def last_error() -> str:
    return load().plb_last_error_string().decode("utf-8", "replace")


def call(name: str, *args) -> None:
    """Invoke an ABI function that returns a status code; raise PlbError with the library's message on failure."""
    rc = getattr(load(), name)(*args)
    if rc != 0:
        raise PlbError(f"{name} failed (rc={rc}): {last_error()}")


def ptr(t) -> int:
    """Device pointer of a torch tensor (None -> NULL); a plain int is passed through (cached arena addresses)."""
    if t is None:
        return 0
    return t if isinstance(t, int) else t.data_ptr()
