"""CPU: the C-ABI library loads and exports every symbol include/plink2_b200.h declares."""
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    txt = open(os.path.join(ROOT, "include", "plink2_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(pl2(?:gpu)?_[a-z0-9_]+)\s*\(", txt)))


def test_header_symbols_exported():
    lib = os.path.join(ROOT, "plink_ng_b200", "libpl2gpu.so")
    assert os.path.exists(lib), "build first: python -c 'import __graft_entry__ as g; g.build()'"
    out = subprocess.run(["nm", "-D", "--defined-only", lib], capture_output=True, text=True, check=True).stdout
    exported = set(re.findall(r"\bT (pl2(?:gpu)?_[a-z0-9_]+)", out))
    declared = _declared()
    assert len(declared) >= 15
    missing = [s for s in declared if s not in exported]
    assert not missing, f"declared in include/plink2_b200.h but not exported: {missing}"


def test_ctypes_binding_covers_header():
    from plink_ng_b200 import capi

    assert sorted(capi.SIGNATURES) == _declared()
    assert capi.lib.pl2gpu_abi_version() >= 1


def test_no_cpu_fallback_without_device():
    import ctypes as C

    from plink_ng_b200 import capi

    if capi.lib.pl2gpu_device_count() > 0:
        pytest.skip("a GPU is visible")
    h = C.c_void_p()
    assert capi.lib.pl2gpu_ctx_create(0, C.byref(h)) == 1
    assert capi.last_error()
