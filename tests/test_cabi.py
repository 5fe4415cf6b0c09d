"""The C-ABI library: builds for sm_100a, loads without a GPU, exports every symbol the header
declares, and fails loudly (no CPU fallback) when there is no device."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def native():
    import importlib.util
    spec = importlib.util.spec_from_file_location("skr_build", os.path.join(ROOT, "scikit-recommender_b200", "build.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    mod.build()
    from skrec_b200 import _native
    return _native


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "skrec_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(skr_[a-z0-9_]+)\s*\(", text)))


def test_every_declared_symbol_is_exported(native):
    L = native.lib()
    declared = _declared_symbols()
    assert len(declared) >= 15
    for name in declared:
        assert hasattr(L, name), name
    assert sorted(native.SYMBOLS) == declared
    assert L.skr_abi_version() == 1


def test_library_is_self_contained(native):
    # plain C ABI: no torch / python / libcuda link-time dependency (cudart is linked statically,
    # the one driver entry point is resolved at run time)
    import subprocess
    out = subprocess.run(["ldd", native.LIB_PATH], capture_output=True, text=True).stdout
    for banned in ("libtorch", "libpython", "libc10", "libcuda.so"):
        assert banned not in out, out


def test_no_device_is_an_error_not_a_fallback(native):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(native.NativeError) as e:
        native.Context(0)
    assert e.value.code == -2 and "no CUDA device" in str(e.value)


def test_product_does_not_import_the_oracle():
    pkg = os.path.join(ROOT, "scikit-recommender_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f
                assert "oracle/" not in src.replace("oracle/ ", ""), f
