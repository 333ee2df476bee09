"""Builds csrc/*.cu into the in-tree shared library libmrb200.so (sm_100a only).

nvcc cross-compiles without a GPU.  The library links the static CUDA runtime and has no
dependency on torch; Python reaches it through ctypes (see _lib.py).
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libmrb200.so")
SOURCES = ["capi.cu", "scene_kernels.cu", "abstract_kernels.cu", "knn_kernels.cu", "knn_tc_kernels.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "--use_fast_math", "-Xcompiler", "-fPIC", "-Xptxas", "-v"]
# --use_fast_math would also turn sincosf/sqrtf/division into approximations; the kernels ask
# for approximate forms explicitly where they are wanted, so keep IEEE defaults elsewhere.
NVCC_FLAGS.remove("--use_fast_math")
if os.environ.get("MRB_DEFS"):  # experiment switches, e.g. MRB_DEFS="-DMRB_TC_E1"
    NVCC_FLAGS += os.environ["MRB_DEFS"].split()


def nvcc() -> str:
    for c in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", shutil.which("nvcc")):
        if c and os.path.exists(c):
            return c
    raise RuntimeError("nvcc not found: the CUDA extension cannot be built (there is no CPU fallback)")


def sources():
    return [os.path.join(CSRC, s) for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", "mrb200.h")]
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB
    objs = []
    build_dir = os.path.join(HERE, "build")
    os.makedirs(build_dir, exist_ok=True)
    log = []
    for src in sources():
        obj = os.path.join(build_dir, os.path.basename(src) + ".o")
        cmd = [nvcc(), *NVCC_FLAGS, "-c", src, "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        log.append(f"$ {' '.join(cmd)}\n{r.stdout}{r.stderr}")
        if r.returncode:
            raise RuntimeError("nvcc failed:\n" + log[-1])
        objs.append(obj)
    cmd = [nvcc(), "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB, *objs]
    r = subprocess.run(cmd, capture_output=True, text=True)
    log.append(f"$ {' '.join(cmd)}\n{r.stdout}{r.stderr}")
    if r.returncode:
        raise RuntimeError("link failed:\n" + log[-1])
    with open(os.path.join(build_dir, "build.log"), "w") as f:
        f.write("\n".join(log))
    if verbose:
        print("\n".join(log))
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
