"""CPU-only checks of the drop-in boundary: the C-ABI library loads and exports every symbol include/hs_gpu.h declares."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "hs_gpu.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(hs_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from hyperspace_b200 import _native

    assert os.path.exists(_native.LIB_PATH), "libhs_gpu.so not built: run __graft_entry__.build()"
    lib = ctypes.CDLL(_native.LIB_PATH)
    declared = _declared_symbols()
    assert len(declared) >= 20
    for sym in declared:
        assert hasattr(lib, sym), f"{sym} declared in hs_gpu.h but not exported"
    assert sorted(_native.EXPORTED_SYMBOLS) == declared
    assert lib.hs_abi_version() == 1


def test_no_cpu_fallback_without_device():
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from hyperspace_b200 import _native

    with pytest.raises(_native.HyperspaceGpuError) as e:
        _native.Context(0)
    assert e.value.code == _native.HS_ENODEVICE
    assert "no CPU fallback" in str(e.value)


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "hyperspace_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".h", ".cuh", ".cc")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.replace("hs_oracle", "oracle") or "oracle" not in re.sub(r"(#|//).*", "", src), \
                    f"{f} references the oracle"
