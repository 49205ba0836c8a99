"""The C-ABI library loads without a GPU, exports every symbol include/*.h declares, and refuses to
compute (loudly) when no CUDA device exists."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    names = set()
    for fn in os.listdir(os.path.join(ROOT, "include")):
        if fn.endswith(".h"):
            src = open(os.path.join(ROOT, "include", fn)).read()
            src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
            names |= set(re.findall(r"\b((?:plbm|c_plbm|c_slbm|c_lw[46]?|c_fvm)_\w+)\s*\(", src))
    return names


def test_every_declared_symbol_is_exported_and_bound():
    from periodic_lbm_b200 import capi

    lib = C.CDLL(capi.LIB_PATH)
    declared = header_symbols()
    assert len(declared) >= 40
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in include/ but not exported"
    # the ctypes stub binds exactly the declared surface
    assert set(capi.SIGNATURES) == declared, set(capi.SIGNATURES) ^ declared


def test_pair_kernel_query_rejects_a_null_handle():
    from periodic_lbm_b200 import capi

    assert capi.lib.plbm_lbm_pair_kernel(None) == -1
    assert b"null grid handle" in capi.lib.plbm_last_error()


def test_library_is_sm100a_only():
    from periodic_lbm_b200 import capi

    out = os.popen(f"cuobjdump -lelf {capi.LIB_PATH} 2>/dev/null").read()
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def test_no_cpu_fallback():
    import periodic_lbm_b200 as p

    if p.device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(p.PlbmError, match="no CPU fallback"):
        p.alloc_grid(16, 16)
    with pytest.raises(p.PlbmError):
        p.SimPlugin().init((8, 8), 1.0, [0.0] * 64, [0.0] * 128)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "periodic_lbm_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h", ".cpp", ".f90")):
                src = open(os.path.join(dirpath, fn), errors="ignore").read()
                assert "oracle" not in src.lower() or fn == "__init__.py" and "oracle" not in src, (dirpath, fn)
