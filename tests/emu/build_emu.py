"""Build tests/emu/libcofdm_emu.so: the kernel sources compiled by g++ against the thread emulator."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "c-ofdm_b200", "csrc")
LIB = os.path.join(HERE, "libcofdm_emu.so")


def build(force=False):
    srcs = [os.path.join(HERE, f) for f in ("emu_kernels.cpp", "cuda_emu.h")] + \
           [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith((".cuh", ".h", ".hpp"))]
    if not force and os.path.exists(LIB) and all(os.path.getmtime(s) <= os.path.getmtime(LIB) for s in srcs):
        return LIB
    subprocess.run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-pthread", "-I", HERE, "-I", CSRC,
                    "-o", LIB, os.path.join(HERE, "emu_kernels.cpp")], check=True)
    return LIB


if __name__ == "__main__":
    print(build(force=True))
