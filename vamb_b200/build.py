"""Build recipe for the in-tree CUDA extension ``vamb_b200/_vk.so`` (sm_100a only).

Plain ``nvcc -shared``: the library exposes the C ABI of include/vamb_b200.h and links
the CUDA runtime statically, so it has no torch / Python dependency.  The built ``.so`` is
git-ignored but travels with the repo snapshot to the GPU box.
"""
from __future__ import annotations

import glob
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_vk.so")
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC,-ffp-contract=off",
    "--expt-relaxed-constexpr",
    "-shared",
]


def sources() -> list:
    return sorted(glob.glob(os.path.join(HERE, "csrc", "*.cu")))


def needs_build() -> bool:
    if not os.path.isfile(OUT):
        return True
    t = os.path.getmtime(OUT)
    deps = sources() + glob.glob(os.path.join(HERE, "csrc", "*.cuh")) + [
        os.path.join(HERE, "..", "include", "vamb_b200.h")
    ]
    return any(os.path.getmtime(p) > t for p in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return OUT
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", OUT] + sources() + ["-lcuda"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed building vamb_b200/_vk.so")
    return OUT


if __name__ == "__main__":
    print(build(force=True, verbose="-v" in sys.argv))
