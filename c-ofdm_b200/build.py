"""Build libcofdm_b200.so (CUDA kernels + C ABI) in-tree with nvcc for sm_100a.

    python c-ofdm_b200/build.py            # or: __graft_entry__.build()

nvcc cross-compiles without a GPU.  The library links the static CUDA runtime only (no cuFFT, no
cuBLAS, no torch), so the same .so serves the ctypes binding, the C++ facade and plain C callers.
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
LIB = os.path.join(HERE, "libcofdm_b200.so")
SOURCES = sorted(os.path.join(HERE, "csrc", f) for f in os.listdir(os.path.join(HERE, "csrc"))
                 if f.endswith((".cu", ".cuh", ".h", ".hpp")))
SOURCES.append(os.path.join(ROOT, "include", "cofdm.h"))


def nvcc_path():
    for c in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if c and os.path.exists(c):
            return c
    raise RuntimeError("nvcc not found")


def is_stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(s) > t for s in SOURCES)


def build_library(force=False, verbose=False, out=None, extra_flags=()):
    """out / extra_flags: an experimental variant beside the product library (A/B runs through COFDM_LIB_PATH)"""
    if out is None and not force and not is_stale():
        return LIB
    cmd = [nvcc_path(), "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-ldl",
           "-Xcompiler", "-fPIC", "-shared", "-o", out or LIB, os.path.join(HERE, "csrc", "cofdm_host.cu")] + list(extra_flags)
    if verbose:
        cmd.insert(1, "-Xptxas")
        cmd.insert(2, "-v")
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + r.stdout + r.stderr)
    if verbose:
        sys.stderr.write(r.stderr)
    return out or LIB


if __name__ == "__main__":
    a = sys.argv[1:]
    out = a[a.index("--out") + 1] if "--out" in a else None
    flags = a[a.index("--flags") + 1].split() if "--flags" in a else ()
    print(build_library(force="--force" in a, verbose="-v" in a, out=out, extra_flags=flags))
