"""The C-ABI shared library loads on a CPU-only box and exports every symbol include/plbert_b200.h declares
(no compute calls here: there is no GPU in the build container)."""
import os
import re

import pytest

from plbert_b200 import _abi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    src = open(os.path.join(ROOT, "include", "plbert_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(plb_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = _abi.load()
    names = declared_functions()
    assert len(names) >= 18
    for n in names:
        assert hasattr(lib, n), f"{n} declared in the header but not exported"
    assert sorted(_abi.SIGNATURES) == names, "ctypes binding table and header drifted apart"


def test_abi_version_and_error_slot():
    lib = _abi.load()
    assert lib.plb_abi_version() == _abi.ABI_VERSION
    # argument validation happens before any CUDA call, so it works without a device and never throws
    rc = lib.plb_gemm_bf16(0, 0, 0, 8, 8, None, 8, None, 8, None, 8, None, None, 0, None, 0, None)
    assert rc == -1 and b"bad shape" in lib.plb_last_error_string()
    rc = lib.plb_attn_fwd_tc(1, 8, 1, 1, 1, 1, 1, 32, 8, 8, 1, 8, 1, None)
    assert rc == -3 and b"head_dim" in lib.plb_last_error_string()
    with pytest.raises(_abi.PlbError):
        _abi.call("plb_ln_fwd", 0, 4, 8, 0, 0, 1e-12, 0, 0, 0)


def test_header_cites_the_reference():
    src = open(os.path.join(ROOT, "include", "plbert_b200.h")).read()
    for needle in ("R/train.py:34-44", "R/train.py:107-131", "R/model.py:15-16", "HF/:71-110", "R/train.py:272"):
        assert needle in src
