"""Build libsphsm_b200.so (the C-ABI CUDA library) in-tree with nvcc for sm_100a.

    python -m sph_sm_monodomain_b200.build [--force]

The .so is git-ignored but travels to the GPU box with the gpurun snapshot.
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libsphsm_b200.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC,-O2,-Wall,-Wno-unused-function",
    "-Xptxas", "-v",
    "--shared", "-cudart", "shared",
]


def _sources():
    srcs = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith((".cu", ".cuh"))]
    srcs.append(os.path.join(ROOT, "include", "sphsm_b200.h"))
    return srcs


HOST = os.path.join(HERE, "host")
DROPIN_INC = os.path.join(ROOT, "include", "dropin")
DROPIN_LIB = os.path.join(HERE, "libsphsm_dropin.so")
HEADLESS = os.path.join(HERE, "sphsm_headless")


def build_dropin(force: bool = False) -> str:
    """g++ -> libsphsm_dropin.so (the reference-compatible C++ class over the C-ABI) and the headless driver that
    replays main.cpp's run protocol.  Host C++11 only, like the reference; links against libsphsm_b200.so."""
    srcs = [os.path.join(HOST, "SPH_SM_monodomain.cpp"), os.path.join(HOST, "headless_main.cpp"),
            os.path.join(ROOT, "include", "sphsm_b200.h")] + [os.path.join(DROPIN_INC, f) for f in os.listdir(DROPIN_INC)]
    outs = [DROPIN_LIB, HEADLESS]
    if not force and all(os.path.exists(o) for o in outs) and min(os.path.getmtime(o) for o in outs) >= max(
            max(os.path.getmtime(x) for x in srcs), os.path.getmtime(LIB)):
        return DROPIN_LIB
    cxx = os.environ.get("CXX", "g++")
    common = [cxx, "-std=c++11", "-O2", "-Wall", "-fPIC", "-I", DROPIN_INC]
    rpath = "-Wl,-rpath,$ORIGIN"
    cmds = [
        common + ["-shared", "-o", DROPIN_LIB, os.path.join(HOST, "SPH_SM_monodomain.cpp"), "-L", HERE, "-lsphsm_b200", rpath],
        common + ["-o", HEADLESS, os.path.join(HOST, "headless_main.cpp"), "-L", HERE, "-lsphsm_dropin", "-lsphsm_b200", rpath],
    ]
    for cmd in cmds:
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            sys.stderr.write(" ".join(cmd) + "\n" + res.stdout + res.stderr)
            raise RuntimeError("g++ failed building the drop-in class")
    return DROPIN_LIB


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(s) > t for s in _sources())


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc] + NVCC_FLAGS + ["-ccbin", "g++", "-o", LIB, os.path.join(CSRC, "sphsm_capi.cu"), "-ldl"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    log = os.path.join(HERE, "build.log")
    with open(log, "w") as fh:
        fh.write(" ".join(cmd) + "\n" + res.stdout + res.stderr)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError(f"nvcc failed (see {log})")
    if verbose:
        print(res.stderr)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
    print(build_dropin(force="--force" in sys.argv))
