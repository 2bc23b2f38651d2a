"""The C-ABI shared library: builds for sm_100a, loads, exports every symbol include/cofdm.h declares,
and fails loudly (never falls back) without a GPU.  No compute calls here."""
import ctypes as C
import os
import re

import pytest

import cofdm_b200 as cb
from conftest import ROOT, has_gpu


@pytest.fixture(scope="module")
def lib():
    cb.build.build_library()
    return cb.load_library()


def test_every_declared_symbol_is_exported(lib):
    hdr = open(os.path.join(ROOT, "include", "cofdm.h")).read()
    names = sorted(set(re.findall(r"\b(cofdm_[a-z0-9_]+)\s*\(", hdr)))
    assert len(names) >= 18
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/cofdm.h but not exported"


def test_library_has_sm100a_code_and_no_fft_library():
    out = os.popen(f"cuobjdump -lelf {cb.LIB_PATH} 2>/dev/null").read()
    assert "sm_100a" in out
    deps = os.popen(f"ldd {cb.LIB_PATH}").read()
    assert "cufft" not in deps and "cublas" not in deps and "torch" not in deps


def test_config_error_is_reported(lib):
    h = C.c_void_p()
    rc = lib.cofdm_create(b"/nonexistent/config.txt", 0, C.byref(h))
    assert rc == -1 and b"Cannot open config file" in lib.cofdm_last_error()


@pytest.mark.skipif(has_gpu(), reason="checks the no-GPU behaviour")
def test_no_gpu_means_loud_failure():
    with pytest.raises(cb.CofdmError, match="no CPU fallback"):
        cb.Modem(os.path.join(ROOT, "config", "config.txt"))


def test_package_never_touches_the_oracle():
    """the product tree must not import, link or load anything under oracle/"""
    pkg = os.path.join(ROOT, "c-ofdm_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp")):
                txt = open(os.path.join(dp, f), errors="ignore").read()
                assert "oracle/" not in txt.replace("no oracle", "") or f == "synth.py" and "from oracle" not in txt, f
                assert "import oracle" not in txt and "from oracle" not in txt and "libcofdm_oracle" not in txt, f


def test_cxx_facade_compiles_against_the_abi(tmp_path):
    """the FRAME_FORM / Modulation look-alikes are plain C++17 over include/cofdm.h"""
    import subprocess
    cb.build.build_library()
    r = subprocess.run(["g++", "-std=c++17", "-O1", f"-I{ROOT}/include", f"-I{ROOT}/c-ofdm_b200/cxx", f"{ROOT}/tests/cxx/main_like.cpp",
                        "-o", str(tmp_path / "main_like"), f"-L{ROOT}/c-ofdm_b200", "-lcofdm_b200"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_txrx_lookalike_with_mac_header_compiles(tmp_path):
    """tx.cpp / rx.cpp look-alike: the drop-in tree now carries mac/mac_frame.hpp (missing from the reference's own tree)"""
    import subprocess
    cb.build.build_library()
    r = subprocess.run(["g++", "-std=c++17", "-O1", f"-I{ROOT}/include", f"-I{ROOT}/c-ofdm_b200/cxx", f"{ROOT}/tests/cxx/txrx_like.cpp",
                        "-o", str(tmp_path / "txrx_like"), f"-L{ROOT}/c-ofdm_b200", "-lcofdm_b200"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_mac_header_reproduces_the_golden_frame(tmp_path):
    """MAC::write on the recorded payload gives the recorded frame (header 01 00 00 00 00 00 7E 57), MAC::read returns it"""
    import subprocess
    import numpy as np
    g = np.load(os.path.join(ROOT, "tests", "golden", "ref_capture.npz"))
    mf = g["mac_frame"].astype(np.uint8)
    src = tmp_path / "mac_check.cpp"
    src.write_text(r'''
#include <cstdio>
#include <fstream>
#include <iterator>
#include "mac/mac_frame.hpp"
int main(int argc, char **argv) {
    std::ifstream f(argv[1], std::ios::binary);
    std::vector<uint8_t> frame((std::istreambuf_iterator<char>(f)), {});
    MAC mac(1, 0, frame.size());
    std::vector<uint8_t> pay(frame.begin() + 8, frame.end());
    auto w = mac.write(pay, 0);
    if (w != frame) { std::printf("write differs\n"); return 1; }
    MAC r(1, 0, frame.size());
    auto back = r.read(frame);
    if (back != pay || !r.checksum_ok() || r.input_tx_id != 1 || r.input_rx_id != 0 || r.input_seq_num != 0) { std::printf("read differs\n"); return 1; }
    frame[20] ^= 1;
    r.read(frame);
    if (r.checksum_ok()) { std::printf("corruption not seen\n"); return 1; }
    auto w2 = mac.write(pay, 0);
    if (w2[4] != 1) { std::printf("seq does not advance\n"); return 1; }
    std::printf("ok %zu\n", mac.payload);
    return 0;
}''')
    exe = tmp_path / "mac_check"
    r = subprocess.run(["g++", "-std=c++17", "-O1", f"-I{ROOT}/c-ofdm_b200/cxx", str(src), "-o", str(exe)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    mf.tofile(tmp_path / "frame.bin")
    r = subprocess.run([str(exe), str(tmp_path / "frame.bin")], capture_output=True, text=True)
    assert r.returncode == 0 and "ok 248" in r.stdout, r.stdout + r.stderr
