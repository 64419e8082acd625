"""Builds ``libparllel_b200.so`` in-tree with nvcc for sm_100a (B200).

The shared library exposes only the C ABI declared in ``include/parllel_b200.h``.
nvcc cross-compiles without a GPU; the built ``.so`` is git-ignored but travels
with the source tree.  ``python -m parllel_b200.build [--force] [--verbose]``.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG_DIR, "csrc")
LIB_PATH = os.path.join(PKG_DIR, "libparllel_b200.so")
SOURCES = ["plb_api.cu", "plb_scan.cu", "plb_stats.cu", "plb_norm.cu", "plb_gather.cu",
           "plb_obs.cu", "plb_pipeline.cu", "plb_exchange.cu", "plb_rng.cu"]
HEADERS = ["plb_common.cuh", "plb_scan_shared.cuh", os.path.join("..", "..", "include", "parllel_b200.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo", "--threads", "0",
    "-Xcompiler", "-fPIC,-fvisibility=hidden,-O3",
    "-Xptxas", "-v",
    "--expt-relaxed-constexpr", "--extended-lambda",
    # shared CUDA runtime: the process (torch) already maps libcudart.so.12, and a statically linked
    # runtime would bake every cudart entry-point name into the artefact
    "-shared", "-cudart", "shared",
    "-Xlinker", "-rpath,/usr/local/cuda/lib64",
    "-Xlinker", "--no-undefined",   # a definition lost in an edit must fail the build, not the first dlopen
]


def _nvcc() -> str:
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found; libparllel_b200.so cannot be built")
    return nvcc


def _sources() -> list[str]:
    return [os.path.join(CSRC, s) for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]


def needs_build() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = _sources() + [os.path.join(CSRC, h) for h in HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.exists(d) and os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB_PATH
    cmd = [_nvcc()] + NVCC_FLAGS + ["-o", LIB_PATH] + _sources()
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or proc.returncode != 0:
        sys.stderr.write(proc.stdout + proc.stderr)
    if proc.returncode != 0:
        raise RuntimeError("nvcc failed building libparllel_b200.so")
    log = os.path.join(PKG_DIR, "csrc", "ptxas.log")
    with open(log, "w") as f:
        f.write(proc.stdout + proc.stderr)
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
