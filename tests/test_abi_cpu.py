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


def test_ctypes_structs_match_header(tmp_path):
    """The argument structs crossing the C ABI: size and every field offset of the ctypes mirrors (mla_b200/_lib.py)
    equal what a C compiler derives from include/mla_b200.h."""
    import shutil
    import subprocess
    from mla_b200 import _lib
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("no C compiler")
    pairs = {"mla_gemm_args": _lib.GemmArgs, "mla_attn_args": _lib.AttnArgs, "mla_mha_args": _lib.MhaArgs,
             "mla_gen_image_args": _lib.GenImageArgs, "mla_gemv_args": _lib.GemvArgs,
             "mla_decode_stack_args": _lib.DecodeStackArgs}
    lines = ['#include <stdio.h>', '#include <stddef.h>', f'#include "{os.path.join(ROOT, "include", "mla_b200.h")}"',
             'int main(void) {']
    for cname, ct in pairs.items():
        lines.append(f'  printf("{cname} size %zu\\n", sizeof({cname}));')
        for fname, _ in ct._fields_:
            lines.append(f'  printf("{cname} {fname} %zu\\n", offsetof({cname}, {fname}));')
    lines += ['  return 0;', '}']
    src = tmp_path / "abi.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "abi"
    r = subprocess.run([gcc, "-std=c11", "-o", str(exe), str(src)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr          # also proves the header is plain C and names every mirrored field
    out = subprocess.run([str(exe)], capture_output=True, text=True).stdout
    want = {}
    for line in out.splitlines():
        s, f, v = line.split()
        want[(s, f)] = int(v)
    for cname, ct in pairs.items():
        assert ctypes.sizeof(ct) == want[(cname, "size")], cname
        for fname, _ in ct._fields_:
            assert getattr(ct, fname).offset == want[(cname, fname)], (cname, fname)
