"""The C-ABI library loads and exports every symbol include/crgpu.h declares (no compute: CPU box)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    src = open(os.path.join(ROOT, "include", "crgpu.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(crgpu_[a-z_0-9]+)\s*\(", src)))


def test_header_declares_the_expected_surface():
    names = declared_functions()
    for must in ("crgpu_create", "crgpu_destroy", "crgpu_qualfilter", "crgpu_align", "crgpu_quantify",
                 "crgpu_align_quantify", "crgpu_last_error", "crgpu_last_timing", "crgpu_int_peak"):
        assert must in names


def test_library_exports_every_declared_symbol():
    from crispresso_b200 import build
    lib = ctypes.CDLL(build.build())
    for name in declared_functions():
        assert hasattr(lib, name), name
    assert lib.crgpu_abi_version() == 5


def test_no_cpu_fallback_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from crispresso_b200 import Context, CrgpuError
    with pytest.raises(CrgpuError):
        Context(0)


def test_struct_layouts_match_the_header():
    from crispresso_b200 import _lib
    assert _lib.ALN_REC.itemsize == 32 and _lib.READ_REC.itemsize == 16
    assert ctypes.sizeof(_lib.QuantParams) == 40
    assert ctypes.sizeof(_lib.PathParams) == 40
    assert ctypes.sizeof(_lib.PathOut) == 8 * 8 + 8 * 8 + 3 * 8 + 8 + 8 + 4 * 8 + 2 * 8 + 4 * 8 + 8 + 8


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "crispresso_b200")
    for dirpath, _d, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in txt.replace("oracle-free", ""), os.path.join(dirpath, f)
