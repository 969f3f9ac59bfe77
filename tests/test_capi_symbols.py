"""The C-ABI library loads and exports every symbol include/metacov_b200.h declares
(no compute calls: this runs without a GPU)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "metacov_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(mcov_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from metacov_b200 import _capi
    lib = ctypes.CDLL(_capi.LIB_PATH)
    names = declared_symbols()
    assert len(names) >= 35
    for name in names:
        assert hasattr(lib, name), "missing export: " + name
    # and the ctypes table binds exactly the declared set
    assert sorted(_capi.SIGNATURES) == names


def test_abi_version_and_struct_sizes():
    from metacov_b200 import _capi
    assert _capi.lib.mcov_abi_version() == 1
    assert ctypes.sizeof(_capi.RegionStats) == 64
    assert ctypes.sizeof(_capi.Filter) == 12
    f = _capi.Filter()
    _capi.lib.mcov_default_filter(ctypes.byref(f))
    assert (f.flag_filter, f.flag_require, f.min_mapq, f.ignore_orphans, f.max_depth) == (0x704, 0, 0, 1, 8000)


def test_no_cpu_fallback_without_gpu():
    """Without a CUDA device the product path must fail loudly, not compute on the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from metacov_b200 import CoverageEngine, McovError
    with pytest.raises(McovError):
        CoverageEngine([100, 200])


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "metacov_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f
                assert "liboracle" not in src, f
