import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import cofdm_b200  # noqa: E402,F401  registers the `c-ofdm_b200/` package as cofdm_b200
from cofdm_b200 import synth  # noqa: E402
from oracle import oracle as oracle_mod  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")
DEFAULT_CONFIG = os.path.join(ROOT, "config", "config.txt")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


def rel_l2(a, b):
    a = np.asarray(a).astype(np.complex128).ravel()
    b = np.asarray(b).astype(np.complex128).ravel()
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


@pytest.fixture(scope="session")
def golden_capture():
    return np.load(os.path.join(GOLDEN, "ref_capture.npz"))


@pytest.fixture(scope="session")
def golden_vectors():
    return np.load(os.path.join(GOLDEN, "ref_vectors.npz"))


@pytest.fixture(scope="session")
def oracle_lib():
    """The C restatement; (re)built from oracle/ if missing or stale."""
    oracle_mod.build("port")
    return oracle_mod


@pytest.fixture(scope="session")
def cfg_dir(tmp_path_factory):
    d = tmp_path_factory.mktemp("cfg")
    paths = {}
    for mt in (1, 2, 4, 6, 8):
        paths[mt] = synth.write_config(str(d / f"config_m{mt}.txt"), modType=mt)
    paths["stream"] = synth.write_config(str(d / "config_stream.txt"), rx_buf_size=10)
    # a small non-default geometry for the generic (any-size) path: the coarse-CFO window width
    # size*(ND+NP)/N/NP = 160*64/128/8 = 10 is an exact integer, as the reference's estimator needs
    paths["small"] = synth.write_config(str(d / "config_small.txt"), fft_size=128, cp_size=32, num_data_subc=56, num_pilot_subc=8,
                                        num_symb=3, pr_sin_len=32, modType=4)
    # BASELINE.json configs[4]: 4096-point, 64-QAM, dense pilots (window width 5120*2048/4096/128 = 20)
    paths["big"] = synth.write_config(str(d / "config_big.txt"), fft_size=4096, cp_size=1024, num_data_subc=1920, num_pilot_subc=128,
                                      num_symb=8, pr_sin_len=128, modType=6)
    return paths


@pytest.fixture(scope="session")
def port(oracle_lib, cfg_dir):
    """dict modType -> Oracle('port')"""
    return {mt: oracle_lib.Oracle("port", cfg_dir[mt]) for mt in (1, 2, 4, 6, 8)}


def has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False
