"""Build recipe for libfdtd_b200.so (nvcc, sm_100a only, in-tree so the .so travels with gpurun snapshots)."""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, "csrc")
# FDTD_B200_LIB: build / load an alternative binary (A/B experiments with build-time switches), default the in-tree one
LIB = os.environ.get("FDTD_B200_LIB") or os.path.join(PKG, "libfdtd_b200.so")
SOURCES = ["fdtd_capi.cu", "nccl_ring.cu", "peer_ring.cu"]
HEADERS = ["fdtd_common.cuh", "sweep_kernels.cuh", "fused_kernel.cuh", "fused_kernel_v2.cuh", "fused_kernel_t2.cuh", "solver.h", os.path.join("..", "..", "include", "fdtd_b200.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-fmad=false",                 # belt and braces: arithmetic already uses __dadd_rn/__dmul_rn (no contraction)
    "-Xcompiler", "-fPIC", "-shared",
    "-Xptxas", "-v",
] + (["-DFDTD_T2_ABLATE"] if os.environ.get("FDTD_T2_ABLATE") else []) \
  + (["-DFDTD_T2_F32_MAGIC=" + os.environ["FDTD_T2_F32_MAGIC"]] if os.environ.get("FDTD_T2_F32_MAGIC") else []) \
  + (["-DFDTD_T2_RINGUP=" + os.environ["FDTD_T2_RINGUP"]] if os.environ.get("FDTD_T2_RINGUP") else [])


def nvcc() -> str:
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found: libfdtd_b200.so cannot be built (there is no CPU fallback)")
    return exe


def stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in SOURCES + HEADERS]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not stale():
        return LIB
    cmd = [nvcc(), *NVCC_FLAGS, "-I", os.path.join(PKG, "..", "include"), "-o", LIB,
           *[os.path.join(CSRC, f) for f in SOURCES], "-ldl"]
    env = dict(os.environ)
    # the image's CC/CXX wrappers lack libgomp specs; use the system compiler as nvcc's host compiler
    cmd[1:1] = ["-ccbin", shutil.which("g++", path="/usr/bin:/bin") or "g++"]
    r = subprocess.run(cmd, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    log = os.path.join(PKG, "build.log")
    with open(log, "w") as fh:
        fh.write(" ".join(cmd) + "\n" + r.stdout)
    if verbose or r.returncode != 0:
        sys.stderr.write(r.stdout)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed (see {log})")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
