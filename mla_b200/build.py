"""Builds mla_b200/lib/libmla_b200.so from mla_b200/csrc/*.cu with nvcc for sm_100a (in-tree, no JIT cache).

`python -m mla_b200.build` or `mla_b200.build.build()`; objects are cached on source mtime so incremental
rebuilds take seconds.  The .so links only the static CUDA runtime: it loads on a CPU-only box (the C-ABI export
test needs that) and fails loudly at the first compute call there (mla_device_check).
"""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

ROOT = Path(__file__).resolve().parent
CSRC = ROOT / "csrc"
OUT_DIR = ROOT / "lib"
OBJ_DIR = OUT_DIR / "obj"
LIB = OUT_DIR / "libmla_b200.so"

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
CFLAGS = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr",
          "-Xptxas", "-v" if os.environ.get("MLA_PTXAS_V") else "-O3"]


def _stale(obj: Path, deps: list[Path]) -> bool:
    if not obj.exists():
        return True
    t = obj.stat().st_mtime
    return any(d.stat().st_mtime > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> Path:
    OBJ_DIR.mkdir(parents=True, exist_ok=True)
    sources = sorted(CSRC.glob("*.cu"))
    headers = sorted(CSRC.glob("*.cuh")) + [ROOT.parent / "include" / "mla_b200.h"]
    jobs = []
    for src in sources:
        obj = OBJ_DIR / (src.stem + ".o")
        if force or _stale(obj, [src] + headers):
            jobs.append((src, obj))

    def compile_one(job):
        src, obj = job
        cmd = [NVCC, *ARCH, *CFLAGS, "-c", str(src), "-o", str(obj)]
        r = subprocess.run(cmd, capture_output=True, text=True)
        return src, r

    if jobs:
        with ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
            for src, r in ex.map(compile_one, jobs):
                if verbose or r.returncode != 0:
                    sys.stderr.write(f"[nvcc] {src.name}\n{r.stdout}{r.stderr}\n")
                if r.returncode != 0:
                    raise RuntimeError(f"nvcc failed on {src}")
    objs = [OBJ_DIR / (s.stem + ".o") for s in sources]
    if force or jobs or not LIB.exists():
        cmd = [NVCC, *ARCH, "-shared", "-cudart", "static", "-o", str(LIB), *map(str, objs)]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError("link failed")
    return LIB


if __name__ == "__main__":
    p = build(force="--force" in sys.argv, verbose=True)
    print(p)
