"""Host-side formats of c-ofdm_b200/apps.py (io/io.hpp dumps, LOG.txt trace, config parser): no GPU needed."""
import numpy as np
import pytest

from cofdm_b200 import apps


def test_dump_formats_round_trip(tmp_path):
    z = (np.arange(10) - 3.5) + 1j * np.linspace(-1, 1, 10)
    apps.write_complex(tmp_path / "c.bin", z)
    raw = np.fromfile(tmp_path / "c.bin", dtype=np.float64)           # exactly what python_code/ofdm.py does
    assert np.array_equal(raw[::2] + 1j * raw[1::2], z)
    assert np.array_equal(apps.read_complex(tmp_path / "c.bin"), z)
    i16 = np.arange(-6, 6, dtype=np.int16).reshape(-1, 2)
    apps.write_complex(tmp_path / "s.bin", i16)                        # data/source.bin: complex<int16_t>
    assert (tmp_path / "s.bin").stat().st_size == i16.size * 2
    assert np.array_equal(apps.read_complex(tmp_path / "s.bin", np.int16), i16)
    apps.write_double(tmp_path / "d.bin", [0.0, 0.93, 0.0])
    assert np.array_equal(np.fromfile(tmp_path / "d.bin", dtype=np.float64), [0.0, 0.93, 0.0])
    (tmp_path / "odd.bin").write_bytes(b"\0" * 24)
    with pytest.raises(ValueError):
        apps.read_complex(tmp_path / "odd.bin")


def test_trace_is_parsed_like_timetrace_py(tmp_path):
    """the trace carries MEASURED stage times (passed in by the caller from cofdm_last_stage_ms), amortised per frame; the
    frames-per-block column comes from the detected positions"""
    stage_ms = {"upload": 0.4, "scan": 0.05, "merge": 0.01, "gather": 0.002, "acquire": 0.02, "demod": 0.03, "d2h": 0.004}
    pos = np.array([300, 7000, 20000, 26000, 41000], dtype=np.int64)       # SDR blocks of 18048 samples: 2 + 2 + 1 frames
    apps.write_trace(tmp_path / "LOG.txt", pos, stage_ms, t_rx=5e-4, t_mac=5e-6, seqs=[7, 8, 9, 10, 11], block_samples=18048)
    rows = []
    for line in open(tmp_path / "LOG.txt"):                           # python_code/timetrace.py parse_log_file
        d = {}
        for part in line.strip().split(" "):
            k, v = part.split(":", 1)
            d[k] = float(v) if ("." in v or "e" in v) else int(v)
        rows.append(d)
    assert [r["ITER"] for r in rows] == [0, 1, 2, 3, 4] and [r["SEQ"] for r in rows] == [7, 8, 9, 10, 11]
    assert [r["FR_IN_BUF"] for r in rows] == [1, 2, 1, 2, 1]
    for key in ("GLOBAL", "T2SIN", "PILOT_SINH", "FREQ_PHASE_SINH", "PFC", "MAC", "DET", "TIME"):
        assert all(key in r for r in rows)
    assert abs(rows[0]["T2SIN"] - 0.05e-3 / 5) < 1e-12 and abs(rows[0]["PILOT_SINH"] - 0.02e-3 / 5) < 1e-12
    assert abs(rows[0]["PFC"] - 0.03e-3 / 5) < 1e-12 and rows[0]["FREQ_PHASE_SINH"] == 0
    assert ["CONVERT" in r for r in rows] == [True, False, True, False, True] and abs(rows[0]["CONVERT"] - 0.4e-3 / 3) < 1e-9
    assert abs(rows[4]["GLOBAL"] - 4 * rows[0]["TIME"]) < 1e-9


def test_config_value_follows_parser_cpp(tmp_path):
    p = tmp_path / "c.txt"
    p.write_text("# comment\n  fft_size = 512  \nT2_sin_level=800\n\nnoequals\n mult\t=\t200\n#mult = 5\n")
    assert apps.config_value(p, "fft_size") == 512 and apps.config_value(p, "T2_sin_level") == 800
    assert apps.config_value(p, "mult") == 200 and apps.config_value(p, "missing") == 0
