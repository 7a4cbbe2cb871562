"""Build libb2rl.so (the C-ABI CUDA library) in-tree with nvcc for sm_100a.

    python -m sac_td3_cudagraphs_pytorch_b200.build [--force] [--verbose]

The .so is git-ignored but travels to the GPU box with the repo snapshot. nvcc cross-compiles
without a GPU, so this runs in the authoring container too.
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
INCLUDE = PKG.parent / "include"
LIB = Path(os.environ.get("B2RL_LIB") or PKG / "libb2rl.so")  # B2RL_LIB: an instrumented build beside the product one
STAMP = PKG / "csrc" / (".build_stamp" if LIB.name == "libb2rl.so" else f".build_stamp_{LIB.stem}")

SOURCES = ["api.cu", "replay.cu", "critic.cu", "actor.cu", "wgrad.cu", "adam.cu", "tc_linear.cu", "wide.cu", "tc_wgrad.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC",
    "-Xptxas", "-v", "-shared", "-cudart", "static",
]


def _nvcc() -> str:
    for c in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if c and Path(c).exists():
            return c
    raise RuntimeError("nvcc not found: libb2rl.so cannot be built (there is no CPU fallback)")


def _digest() -> str:
    h = hashlib.sha256()
    for f in sorted(list(CSRC.glob("*.cu")) + list(CSRC.glob("*.cuh")) + [INCLUDE / "b2rl.h", Path(__file__)]):
        h.update(f.name.encode())
        h.update(f.read_bytes())
    h.update(os.environ.get("B2RL_EXTRA_NVCC_FLAGS", "").encode())
    return h.hexdigest()


def build_lib(force: bool = False, verbose: bool = False) -> Path:
    digest = _digest()
    if not force and LIB.exists() and STAMP.exists() and STAMP.read_text().strip() == digest:
        return LIB
    extra = os.environ.get("B2RL_EXTRA_NVCC_FLAGS", "").split()
    cmd = [_nvcc(), *NVCC_FLAGS, *extra, "-I", str(INCLUDE), "-o", str(LIB)] + [str(CSRC / s) for s in SOURCES]
    proc = subprocess.run(cmd, capture_output=True, text=True)
    log = proc.stdout + proc.stderr
    (CSRC / ".build_log.txt").write_text(" ".join(cmd) + "\n" + log)
    if proc.returncode != 0:
        sys.stderr.write(log)
        raise RuntimeError("nvcc failed building libb2rl.so")
    if verbose:
        print(log)
    STAMP.write_text(digest)
    return LIB


if __name__ == "__main__":
    p = build_lib(force="--force" in sys.argv, verbose="--verbose" in sys.argv)
    print(p)
