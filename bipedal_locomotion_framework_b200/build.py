"""In-tree build of the native pieces (no JIT cache: the .so files travel with the snapshot).

  lib/libblf_ccm.so        CUDA kernels + C ABI (nvcc, sm_100a only)
  lib/libblf_contact.so    C++17 host facade over the C ABI (g++)            [build_cpp]
  lib/blf_cpp_tests        the reference's Catch2 sections restated          [build_cpp]
"""
from __future__ import annotations

import os
import shutil
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(_HERE)
LIB_DIR = os.path.join(_HERE, "lib")
CUDA_LIB = os.path.join(LIB_DIR, "libblf_ccm.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC,-fvisibility=hidden", "-shared",
]


def _newer(target: str, sources: list[str]) -> bool:
    if not os.path.exists(target):
        return False
    t = os.path.getmtime(target)
    return all(os.path.getmtime(s) <= t for s in sources)


def _nvcc() -> str:
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; cannot build the sm_100a library")


def cuda_sources() -> list[str]:
    c = os.path.join(_HERE, "csrc")
    return [os.path.join(c, f) for f in sorted(os.listdir(c))] + \
        [os.path.join(ROOT, "include", "blf_ccm.h")]


def build_cuda(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(LIB_DIR, exist_ok=True)
    if not force and _newer(CUDA_LIB, cuda_sources()):
        return CUDA_LIB
    cmd = [_nvcc(), *NVCC_FLAGS, "-o", CUDA_LIB, os.path.join(_HERE, "csrc", "ccm_capi.cu"),
           "-lcudart", "-ldl"]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + r.stdout)
    if verbose:
        print(r.stdout)
    return CUDA_LIB


if __name__ == "__main__":
    import sys
    print(build_cuda(force="--force" in sys.argv, verbose="-v" in sys.argv))
