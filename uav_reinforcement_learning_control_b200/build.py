"""In-tree build of libquadsim.so for sm_100a (explicit nvcc, no JIT cache).

The .so is git-ignored but travels to the GPU box with the gpurun snapshot, so it must be
built HERE (nvcc cross-compiles without a GPU).  ``build_library`` is idempotent: it
rebuilds only when a source is newer than the library.
"""
from __future__ import annotations

import os
import shutil
import subprocess

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG_DIR, "csrc")
ROOT = os.path.dirname(PKG_DIR)
LIB_PATH = os.path.join(CSRC, "libquadsim.so")

SOURCES = ["quadsim.cu"]
HEADERS = ["qs_math.cuh", "qs_pack2.cuh", "qs_step2.cuh", "qs_philox.cuh", "qs_dynamics.cuh", "qs_env.cuh", "qs_kernels.cuh", "qs_rollout.cuh",
           "qs_rollout_tc.cuh", "qs_ppo.cuh", "qs_ppo_generic.cuh", "qs_traj.cuh"]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "--expt-relaxed-constexpr", "-Xcompiler", "-fPIC", "-shared",
]


def find_nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: libquadsim.so cannot be built (there is no CPU fallback)")


def _deps():
    return ([os.path.join(CSRC, f) for f in SOURCES + HEADERS] +
            [os.path.join(ROOT, "include", "quadsim_abi.h")])


def needs_build() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    return any(os.path.getmtime(d) > t for d in _deps())


def build_library(force: bool = False, verbose: bool = False, out: str | None = None, defines=()) -> str:
    """`out` / `defines` build a tuning variant next to the product library (used for A/B runs)."""
    if out is None and not force and not needs_build():
        return LIB_PATH
    out = out or LIB_PATH
    cmd = [find_nvcc()] + NVCC_FLAGS + [f"-D{d}" for d in defines] + (["-Xptxas", "-v"] if verbose else []) + \
          ["-o", out] + [os.path.join(CSRC, s) for s in SOURCES]
    res = subprocess.run(cmd, cwd=CSRC, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
    if verbose:
        print(res.stderr)
    return out


if __name__ == "__main__":
    import sys
    print(build_library(force="--force" in sys.argv, verbose="-v" in sys.argv))
