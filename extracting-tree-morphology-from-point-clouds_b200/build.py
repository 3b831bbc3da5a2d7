"""Builds ``csrc/*.cu`` into ``libtreemorph_nn.so`` (the C-ABI library of include/treemorph_nn.h).

sm_100a only, in tree (the built .so travels to the GPU box with the repo snapshot):

    nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -shared -Xcompiler -fPIC ...

No fast-math flags: the distance arithmetic relies on IEEE division / square root, no FTZ and no
FMA contraction of the ``__f*_rn`` intrinsics (see csrc/tm_eval.cuh).
"""
from __future__ import annotations

import glob
import os
import shutil
import subprocess

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG_DIR, "csrc")
LIB_PATH = os.path.join(PKG_DIR, "libtreemorph_nn.so")
SOURCES = ["tm_api.cu", "tm_brute.cu", "tm_grid.cu", "tm_small.cu", "tm_bvh.cu", "tm_knn.cu", "tm_noise.cu", "tm_comm.cu"]

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-shared", "-Xcompiler", "-fPIC", "-Xcompiler", "-pthread", "--threads", "8", "-ldl"]


def find_nvcc() -> str | None:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    return None


def is_stale() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = glob.glob(os.path.join(CSRC, "*.cu")) + glob.glob(os.path.join(CSRC, "*.cuh"))
    deps.append(os.path.join(os.path.dirname(PKG_DIR), "include", "treemorph_nn.h"))
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile if sources are newer than the library.  Returns the library path."""
    if not force and not is_stale():
        return LIB_PATH
    nvcc = find_nvcc()
    if nvcc is None:
        raise RuntimeError("nvcc not found: cannot build libtreemorph_nn.so")
    env = dict(os.environ)
    env.pop("CC", None)          # the image exports a gcc wrapper that nvcc must not pick up as host compiler
    env.pop("CXX", None)
    cmd = [nvcc, *NVCC_FLAGS, "-o", LIB_PATH + ".tmp", *[os.path.join(CSRC, s) for s in SOURCES]]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, env=env)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout)
    os.replace(LIB_PATH + ".tmp", LIB_PATH)
    if verbose:
        print(res.stdout)
    return LIB_PATH


if __name__ == "__main__":
    import sys
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
