"""In-tree build of the CUDA library (sm_100a only).  nvcc cross-compiles without a GPU.

    python -m eigen_git_mirror_b200.build          # or __graft_entry__.build()

Output: eigen-git-mirror_b200/lib/libb200sparse.so (git-ignored; travels to the GPU box with the snapshot).
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "lib", "libb200sparse.so")
SOURCES = ["b200sparse.cu", "plan.cpp", "factors.cpp"]
HEADERS = ["kernels.cuh", "kernels_multi.cuh", "kernels_krylov.cuh", "krylov.inc", "precond.inc", "kernels_tri.cuh", "factors.h", "device_state.h", "plan.h", os.path.join("..", "..", "include", "b200sparse.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared"]


def nvcc_path() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", shutil.which("nvcc")):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: the product is CUDA-only and cannot be built without it")


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in SOURCES + HEADERS]
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB
    os.makedirs(os.path.dirname(LIB), exist_ok=True)
    cmd = [nvcc_path(), *NVCC_FLAGS, "-o", LIB, *[os.path.join(CSRC, s) for s in SOURCES]]
    if verbose:
        cmd.insert(1, "-Xptxas")
        cmd.insert(2, "-v")
    env = dict(os.environ)
    env.pop("CXX", None)  # the image's CXX wrapper lacks libgomp specs; nvcc finds /usr/bin/g++ itself
    env.pop("CC", None)
    res = subprocess.run(cmd, capture_output=True, text=True, env=env)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + res.stdout + res.stderr)
    if verbose:
        sys.stderr.write(res.stderr)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
