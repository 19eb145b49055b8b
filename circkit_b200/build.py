"""Builds libcirckit_b200.so in-tree with nvcc for sm_100a (no GPU needed: nvcc cross-compiles)."""
from __future__ import annotations

import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libcirckit_b200.so")
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-shared", "-Xcompiler", "-fPIC"]


def _nvcc() -> str:
    for c in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if c and (os.path.isabs(c) and os.path.exists(c) or not os.path.isabs(c)):
            return c
    return "nvcc"


def sources() -> list[str]:
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh", ".cpp")))


def stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    hdr = os.path.join(os.path.dirname(HERE), "include", "circkit_b200.h")
    return any(os.path.getmtime(s) > t for s in sources() + [hdr])


def build(force: bool = False, verbose: bool = False) -> str:
    if force or stale():
        env = dict(os.environ)
        env.pop("CC", None); env.pop("CXX", None)      # the image's CC wrapper lacks some specs
        # the host packer (CK_F_PACKED_IN) is plain C++ with run-time-dispatched AVX2 / BMI2 paths: g++, then linked in
        obj = os.path.join(HERE, "ck_host_pack.o")
        subprocess.run(["g++", "-O3", "-std=c++17", "-fPIC", "-pthread", "-c", "-o", obj, os.path.join(CSRC, "ck_host_pack.cpp")],
                       check=True, env=env)
        cmd = [_nvcc(), *NVCC_FLAGS, "-o", LIB, os.path.join(CSRC, "ck_lib.cu"), obj, "-Xcompiler", "-pthread"]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        subprocess.run(cmd, check=True, env=env)
    return LIB


if __name__ == "__main__":
    print(build(force=True, verbose=True))
