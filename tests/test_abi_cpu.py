"""CPU checks of the boundary: the C-ABI library builds, loads, exports every symbol include/mla_b200.h declares, and
refuses to compute without an sm_100 device (no silent fallback)."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from mla_b200 import build, _lib
    build.build()
    return _lib.lib()


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "mla_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(mla_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_exported(lib):
    syms = declared_symbols()
    assert len(syms) >= 30
    missing = [s for s in syms if not hasattr(lib, s)]
    assert not missing, f"declared in include/mla_b200.h but not exported: {missing}"


def test_version_and_error_slot(lib):
    assert b"sm_100a" in lib.mla_version()
    assert isinstance(lib.mla_last_error(), bytes)


@pytest.mark.skipif(torch.cuda.is_available(), reason="CPU-only behaviour")
def test_no_cpu_fallback(lib):
    from mla_b200 import _lib, ops
    assert lib.mla_device_check() == -3          # MLA_ERR_DEVICE
    g = _lib.GemmArgs()
    assert lib.mla_gemm_bf16(ctypes.byref(g), None) == -3
    with pytest.raises(_lib.MlaError):
        ops.gemm(torch.zeros(8, 8, dtype=torch.bfloat16), torch.zeros(8, 8, dtype=torch.bfloat16))
    with pytest.raises(_lib.MlaError):
        ops.rmsnorm_fwd(torch.zeros(2, 8, dtype=torch.bfloat16), torch.ones(8, dtype=torch.bfloat16), 1e-5)


def test_product_never_imports_oracle():
    """The oracle is test infrastructure: nothing under mla_b200/ may import it."""
    pkg = os.path.join(ROOT, "mla_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), os.path.join(dirpath, f)
