"""The C-ABI library loads on a CPU-only box and exports every entry point include/cape_msda.h declares.
No compute calls here (those are the `-m gpu` tests)."""
import ctypes
import os
import re
import subprocess

import pytest

from tests.conftest import REPO

HEADER = os.path.join(REPO, "include", "cape_msda.h")


def _declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"CAPE_API\s+[\w\s\*]+?\b(cape_\w+)\s*\(", text)))


@pytest.fixture(scope="module")
def lib_path():
    import cape_b200
    from cape_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        import importlib.util
        spec = importlib.util.spec_from_file_location("cape_build", os.path.join(os.path.dirname(_lib.LIB_PATH), "build.py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        mod.build()
    return _lib.LIB_PATH


def test_header_declares_expected_entry_points():
    syms = _declared_symbols()
    assert {"cape_abi_version", "cape_last_error", "cape_msda_forward", "cape_msda_backward", "cape_msda_decode",
            "cape_msda_forward_backward_host", "cape_msda_host_workspace_bytes", "cape_launch_count"} <= set(syms)


def test_library_exports_every_declared_symbol(lib_path):
    lib = ctypes.CDLL(lib_path)
    for name in _declared_symbols():
        assert hasattr(lib, name), f"{name} declared in include/cape_msda.h but not exported"
    from cape_b200 import _lib
    assert set(_lib.EXPORTED_SYMBOLS) == set(_declared_symbols())


def test_abi_version_and_error_codes_without_gpu(lib_path):
    from cape_b200 import _lib
    lib = _lib.load()
    assert lib.cape_abi_version() == 1
    assert lib.cape_launch_count() >= 0
    # argument validation happens before any CUDA call
    assert lib.cape_msda_forward(None, None, None, None, None, None, None, 0, 0, None) == -3       # NULL dims
    assert b"dims" in lib.cape_last_error()
    bad = _lib.Dims(1, 4, 1, 30, 1, 1, 1)                                                            # D % 4 != 0
    assert lib.cape_msda_forward(None, None, None, None, None, None, ctypes.byref(bad), 0, 0, None) == -1
    bad = _lib.Dims(1, 4, 1, 32, 1, 9, 4)                                                            # L > 8
    assert lib.cape_msda_backward(None, None, None, None, None, None, None, None, None, ctypes.byref(bad),
                                  0, 0, 0, None) == -1
    ok = _lib.Dims(1, 4, 1, 32, 1, 1, 4)
    assert lib.cape_msda_forward(None, None, None, None, None, None, ctypes.byref(ok), 7, 0, None) == -2   # dtype
    assert lib.cape_msda_forward(None, None, None, None, None, None, ctypes.byref(ok), 1, 2, None) == -2   # bf16 + f16 aux
    assert lib.cape_msda_host_workspace_bytes(ctypes.byref(ok), 1) > lib.cape_msda_host_workspace_bytes(ctypes.byref(ok), 0) > 0


def test_only_sm100a_sass_is_embedded(lib_path):
    cuobjdump = "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    out = subprocess.run([cuobjdump, "-lelf", lib_path], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


def test_product_package_does_not_import_the_oracle():
    """oracle/ is test infrastructure: no product source may reference it."""
    pkg = os.path.join(REPO, "category-agnostic-pose-estimation_b200")
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(root, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f
                assert "msda_oracle" not in text, f
