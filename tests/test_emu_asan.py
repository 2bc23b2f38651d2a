"""compute-sanitizer is closed on the GPU pool, so the bounds check of the kernel source is done here: the
kernels compiled for the CPU thread emulator with -fsanitize=address, running a representative subset of
tests/test_emu_kernels.py in a child process (dynamic shared memory and every global buffer are heap blocks,
so an out-of-bounds access of either is reported)."""
import os
import subprocess
import sys

import pytest

from conftest import ROOT


def test_emulated_kernels_are_asan_clean(tmp_path):
    asan = subprocess.run(["gcc", "-print-file-name=libasan.so"], capture_output=True, text=True).stdout.strip()
    if not asan or not os.path.exists(asan):
        pytest.skip("libasan not available")
    emu = os.path.join(ROOT, "tests", "emu")
    lib = tmp_path / "libcofdm_emu_asan.so"
    subprocess.run(["g++", "-O1", "-g", "-std=c++17", "-fPIC", "-shared", "-pthread", "-fsanitize=address", "-fno-omit-frame-pointer",
                    "-I", emu, "-I", os.path.join(ROOT, "c-ofdm_b200", "csrc"), "-o", str(lib), os.path.join(emu, "emu_kernels.cpp")], check=True)
    env = dict(os.environ, LD_PRELOAD=asan, ASAN_OPTIONS="detect_leaks=0:abort_on_error=1", COFDM_EMU_LIB=str(lib))
    r = subprocess.run([sys.executable, "-m", "pytest", "-x", "-q", "-p", "no:cacheprovider", os.path.join(ROOT, "tests", "test_emu_kernels.py"), "-k",
                        "tx or golden_vectors or loopback or sync_kernels or read_syncless or generic or production"],
                       env=env, capture_output=True, text=True, cwd=ROOT, timeout=900)
    assert r.returncode == 0 and "AddressSanitizer" not in r.stderr, r.stdout[-2000:] + r.stderr[-2000:]
