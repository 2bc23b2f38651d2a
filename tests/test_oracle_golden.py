"""Pin the oracle: the C restatement (oracle/cofdm_oracle.c) against
  (1) the reference's own recorded artefacts (tests/golden/ref_capture.npz  <- reference data/*.bin),
  (2) outputs of the UNMODIFIED reference sources compiled in the build container
      (tests/golden/ref_vectors.npz <- oracle/_ref/libcofdm_ref.so via tests/golden/make_golden.py),
  (3) live, the compiled reference itself when oracle/_ref is present (build container only).
Known answers from SURVEY.md section 4 / appendix A are asserted literally."""
import numpy as np
import pytest

from conftest import rel_l2
from cofdm_b200 import synth


def cplx(i16):
    return i16[..., 0].astype(np.float64) + 1j * i16[..., 1].astype(np.float64)


def test_sizes_and_constants(port):
    s = port[4].sizes
    assert (s.usefull_size, s.output_size, s.t2sin_size, s.preamble_size, s.message_size) == (1024, 6016, 256, 640, 5120)
    c = port[4].constants()
    assert list(c["preamble_bytes"][:8]) == [95, 203, 243, 46, 187, 199, 153, 152]          # mt19937(42) >> 24
    n = np.arange(256)
    tone = 0.5 * np.exp(2j * np.pi * 17 * n / 256) + 0.5 * np.exp(2j * np.pi * 51 * n / 256)
    assert np.abs(c["t2sin_tone"] - tone).max() < 1e-13
    assert np.nonzero(c["t2_mask"])[0].tolist() == list(range(12, 23)) + list(range(46, 57))
    assert abs(np.linalg.norm(c["matched"]) - 1) < 1e-14


def test_constellations_known_answers(port):
    o = port[4]
    np.testing.assert_allclose(o.mod(1, [0b10000000])[:2], [(1 + 1j) / np.sqrt(2), -(1 + 1j) / np.sqrt(2)], atol=1e-15)
    np.testing.assert_array_equal(o.mod(2, [0b00011011]), [-1 - 1j, 1 - 1j, -1 + 1j, 1 + 1j])
    pts16 = o.mod(4, [(s << 4) | s for s in range(16)])[::2]
    for s in range(16):
        assert abs(pts16[s] - complex(-1 + 2 * (s % 4) / 3, -1 + 2 * (s // 4) / 3)) < 1e-15
    np.testing.assert_array_equal(o.bit_stream_converter(4, 8, [0xEF, 0xBB, 0xBF, 0x54]), [14, 15, 11, 11, 11, 15, 5, 4])
    np.testing.assert_array_equal(o.bit_stream_converter(6, 8, [0xEF, 0xBB, 0xBF]), [59, 59, 46, 63])
    np.testing.assert_array_equal(o.bit_stream_converter(1, 8, [0xA5]), [1, 0, 1, 0, 0, 1, 0, 1])
    # demap ties go to the upper level; clamping (modulation.cpp:68-78)
    b, clamped = o.demod(4, np.array([-2 / 3 + 0j, 2 / 3 - 1.7j, 0.6666 + 0.3333j, 1e-17 - 1e-17j] * 2))
    assert [b[0] >> 4, b[0] & 15, b[1] >> 4, b[1] & 15] == [9, 3, 10, 10]
    assert clamped[1] == 2 / 3 - 1j


def test_tx_matches_reference_source_bin(oracle_lib, cfg_dir, golden_capture):
    o = oracle_lib.Oracle("port", cfg_dir[1])                    # the recorded run used BPSK
    frame, q = o.tx(golden_capture["mac_frame"])
    assert np.array_equal(q, golden_capture["source_i16"])       # bit exact, 12032 int16
    assert q[:8].tolist() == [200, 0, 122, 135, -13, 133, -50, 37]


def test_rx_chain_matches_reference_dumps(oracle_lib, cfg_dir, golden_capture):
    o = oracle_lib.Oracle("port", cfg_dir[1])
    cap = cplx(golden_capture["capture_i16"])
    corr = o.t2sin_corr(cap)
    assert np.abs(corr - golden_capture["t2_sin_corr"]).max() < 1e-14
    assert np.nonzero(corr)[0].tolist() == [42, 74]
    t2 = o.find_t2sin(cap, 0)
    pr = o.find_preamble(cap, t2)
    assert (t2, pr) == (10752, 11039)
    r = o.rx_aligned(cap[pr + 1: pr + 1 + 5760])
    assert r["scal"][0] == -19 / 5120
    assert np.abs(r["chan"] - golden_capture["phases"]).max() < 1e-14
    assert rel_l2(r["constell"], golden_capture["constell"]) < 1e-13
    assert np.array_equal(r["bytes"], golden_capture["mac_frame"])
    assert np.array_equal(r["bytes"][8:], golden_capture["data_txt"])
    # second copy of the frame in the capture, found by the rx.cpp-style loop
    pos, by = o.rx_stream(golden_capture["capture_i16"][:240640])
    assert pos.tolist() == [11040, 19302]
    assert all(np.array_equal(b, golden_capture["mac_frame"]) for b in by)


@pytest.mark.parametrize("mt", [1, 2, 4, 6, 8])
def test_port_matches_compiled_reference_vectors(port, golden_vectors, mt):
    g, o = golden_vectors, port[mt]
    s = o.sizes
    for i, pay in enumerate(g[f"m{mt}_payload"]):
        frame, q = o.tx(pay)
        assert np.array_equal(q.reshape(-1, 2), g[f"m{mt}_tx_i16"][i])
        if i == 0:
            assert np.array_equal(frame, g[f"m{mt}_tx_frame0"])
    for i, rec in enumerate(cplx(g[f"m{mt}_rx_in_i16"])):
        r = o.rx_aligned(rec)
        for k in ("scal", "chan", "constell", "bytes") + (("synced", "grid") if mt == 4 else ()):
            assert np.array_equal(r[k], g[f"m{mt}_rx_{k}"][i]), (mt, i, k)
    assert np.array_equal(o.chan_char(cplx(g[f"m{mt}_rx_in_i16"])[0][: s.preamble_size]), g[f"m{mt}_chan_char"])
    b, restored = o.read(o.tx(g[f"m{mt}_payload"][1])[0])
    assert np.array_equal(b, g[f"m{mt}_read_bytes"]) and np.array_equal(restored, g[f"m{mt}_read_restored"])
    assert np.array_equal(o.mod(mt, g[f"m{mt}_mod_in"]), g[f"m{mt}_mod_out"])
    db, clamped = o.demod(mt, g[f"m{mt}_demod_in"])
    assert np.array_equal(db, g[f"m{mt}_demod_out"]) and np.array_equal(clamped, g[f"m{mt}_demod_clamped"])


def test_port_sync_and_stream_match_compiled_reference(oracle_lib, cfg_dir, golden_vectors):
    g = golden_vectors
    o = oracle_lib.Oracle("port", cfg_dir["stream"])
    cap = cplx(g["sync_capture_i16"])
    assert np.array_equal(o.t2sin_corr(cap), g["sync_t2corr"])
    assert [o.find_t2sin(cap, int(st)) for st in g["sync_find_t2_starts"]] == g["sync_find_t2"].tolist()
    assert [o.find_preamble(cap, int(p)) for p in g["sync_pre_starts"]] == g["sync_find_pre"].tolist()
    for p, ref in zip(g["sync_pre_starts"], g["sync_find_corr"]):
        assert np.array_equal(o.find_corr(cap, int(p)), ref)
    pos, by = o.rx_stream(g["sync_capture_i16"])
    assert pos.tolist() == g["sync_stream_pos"].tolist()
    assert np.array_equal(by, g["sync_stream_bytes"])
    assert np.array_equal(by, g["sync_payload"])                 # every frame of the capture decodes


def test_port_vs_compiled_reference_live(oracle_lib, cfg_dir):
    """Build container only: fresh random inputs through both checkers must agree bit for bit."""
    if not oracle_lib.available("reference"):
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    rng = np.random.default_rng(2024)
    for mt in (2, 4, 6):
        R, P = oracle_lib.Oracle("reference", cfg_dir[mt]), oracle_lib.Oracle("port", cfg_dir[mt])
        s = R.sizes
        pay = rng.integers(0, 256, s.usefull_size, dtype=np.uint8)
        fr, qr = R.tx(pay)
        fp, qp = P.tx(pay)
        assert np.array_equal(fr, fp) and np.array_equal(qr, qp)
        rx = synth.channel(qr.reshape(-1, 2), seed=mt, cfo=0.0013, phase=0.3, taps=(1, 0.1j), noise_sigma=2.0)
        rec = rx[s.t2sin_size: s.t2sin_size + s.preamble_size + s.message_size]
        a, b = R.rx_aligned(rec), P.rx_aligned(rec)
        for k in a:
            assert np.array_equal(a[k], b[k]), (mt, k)


def test_config_errors(oracle_lib, tmp_path):
    with pytest.raises(RuntimeError, match="Cannot open config file"):
        oracle_lib.Oracle("port", str(tmp_path / "missing.txt"))
