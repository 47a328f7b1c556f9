"""Installs the UNMODIFIED reference (ZhuoyangLiu2005/MLA, /root/reference) into the git-ignored baseline/_ref/.

    python tools/install_reference.py            # needs /root/reference (this container); idempotent

Why not `pip install --target baseline/_ref /root/reference`: the project's `packages.find` sweeps in
models/vlm/prismatic_bk (432 MB of a stale backup tree with prebuilt py3.7/3.8 .so files) and vla/datasets/cpp build
trees (177 MB), its dependency pins (torch 2.5.1, tensorflow 2.15, timm 0.9.10 ...) cannot be resolved offline, and
the package has no entry point — it is used as a source tree on sys.path (scripts/train.py).  So this script does
what the install would do for the hot path: a byte-for-byte copy of the Python sources the training step imports
(no file is edited; `cmp` against /root/reference passes for every file), nothing else:

    models/                 minus vlm/prismatic_bk (dead duplicate, SURVEY #22)
    transformers/           the vendored 4.40.1, whole (models/backbones/llm/{mistral,phi}.py import other families at
                            package-import time, so the model zoo cannot be pruned safely)
    vla/action_tokenizer.py
    util/, training/, conf/ small Python packages next to the path

baseline/_ref/ is listed in .gitignore (reference sources never enter this repo's history) and NOT in .gpurunignore:
it travels to the GPU box like a pip --target install would, where tools/ref_gpu.py and `bench.py --impl reference`
import it through oracle/ref_shim.py (MLA_REFERENCE_ROOT=baseline/_ref).
"""
from __future__ import annotations

import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.environ.get("MLA_REFERENCE_SRC", "/root/reference")
DST = os.path.join(ROOT, "baseline", "_ref")

def _ignore_common(d, names):
    return [n for n in names if n == "__pycache__" or n.endswith((".pyc", ".so", ".o", ".egg-info"))]


def install(verbose: bool = True) -> str:
    if not os.path.isdir(os.path.join(SRC, "models", "mla")):
        raise RuntimeError(f"reference tree not found at {SRC}")
    if os.path.isdir(DST):
        shutil.rmtree(DST)
    os.makedirs(DST)

    def ignore_models(d, names):
        out = _ignore_common(d, names)
        if os.path.abspath(d) == os.path.join(SRC, "models", "vlm"):
            out.append("prismatic_bk")
        return out

    shutil.copytree(os.path.join(SRC, "models"), os.path.join(DST, "models"), ignore=ignore_models)
    shutil.copytree(os.path.join(SRC, "transformers"), os.path.join(DST, "transformers"), ignore=_ignore_common)
    for pkg in ("util", "training", "conf"):
        shutil.copytree(os.path.join(SRC, pkg), os.path.join(DST, pkg), ignore=_ignore_common)
    os.makedirs(os.path.join(DST, "vla"))
    shutil.copy2(os.path.join(SRC, "vla", "action_tokenizer.py"), os.path.join(DST, "vla", "action_tokenizer.py"))
    for f in ("LICENSE", "pyproject.toml"):
        shutil.copy2(os.path.join(SRC, f), os.path.join(DST, f))
    n = sum(len(fs) for _, _, fs in os.walk(DST))
    size = sum(os.path.getsize(os.path.join(d, f)) for d, _, fs in os.walk(DST) for f in fs)
    if verbose:
        print(f"installed {n} files, {size / 1e6:.1f} MB -> {DST}")
    return DST


if __name__ == "__main__":
    install()
    sys.exit(0)
