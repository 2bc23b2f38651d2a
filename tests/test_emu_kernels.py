"""The real kernel source (c-ofdm_b200/csrc/*.cuh) executed by the CPU thread emulator and checked
against the oracle: index maps, exchange layouts, bit packing, arithmetic.  Small sizes only."""
import numpy as np
import pytest

import parity_checks as pc
from emu_backend import EmuModem


@pytest.fixture(scope="module")
def emu(cfg_dir, port):
    ms = {mt: EmuModem(cfg_dir[mt], port[mt].sizes) for mt in (1, 2, 4, 6, 8)}
    yield ms
    for m in ms.values():
        m.close()


@pytest.mark.parametrize("mt", [1, 2, 4, 6, 8])
def test_mod_demod(emu, port, mt):
    pc.check_mod_demod(emu[mt], port[mt], mt)


@pytest.mark.parametrize("mt", [1, 4, 8])
def test_tx(emu, port, mt):
    st = pc.check_tx(emu[mt], port[mt], n_frames=2)
    assert st["rel_l2"] < 5e-7


def test_tx_register_store_variant(emu, port):
    """the non-bulk output stage (used when the frame buffer is not 16-byte aligned) gives the same frames"""
    pay = pc.synth.payloads(3, port[4].sizes.usefull_size, seed=9)
    a = emu[4].tx_batch(pay, 1)
    c = emu[4].tx_batch(pay, 0)
    emu[4].set_tx_bulk(0)
    try:
        b = emu[4].tx_batch(pay, 1)
        d = emu[4].tx_batch(pay, 0)
    finally:
        emu[4].set_tx_bulk(1)
    assert np.array_equal(a, b) and np.array_equal(c, d)


@pytest.mark.parametrize("mt,fmt", [(4, "i16"), (4, "cf32"), (2, "i16"), (6, "cf32"), (1, "i16"), (8, "i16")])
def test_rx_fused_against_oracle(emu, port, mt, fmt):
    pay, rec = pc.impaired_records(port[mt], 3, seed=10 * mt)
    st = pc.check_rx_against_oracle(emu[mt], port[mt], rec, fmt)
    assert st["shift_mismatch"] == 0
    assert max(st["synced"], st["grid"], st["constell"]) < 2e-6


def test_rx_fused_golden_vectors(emu, port, golden_vectors):
    """committed outputs of the compiled reference (default modType) as the expected values"""
    g = golden_vectors
    want = [{k: g[f"m4_rx_{k}"][i] for k in ("scal", "synced", "grid", "chan", "constell", "bytes")} for i in range(2)]
    st = pc.check_rx_against_oracle(emu[4], port[4], g["m4_rx_in_i16"], "i16", want=want)
    assert st["shift_mismatch"] == 0 and st["differing"] == 0


def test_rx_non_tma_path_is_identical(emu, port):
    pay, rec = pc.impaired_records(port[4], 2, seed=3)
    x = pc.cplx(rec).astype(np.complex64)
    a, _ = emu[4].rx_aligned_batch(x)
    emu[4].use_tma = 0
    b, _ = emu[4].rx_aligned_batch(x)
    emu[4].use_tma = 1
    assert np.array_equal(a, b)
    # int16 wire records: bulk-copied raw and widened when read, or widened while loading -- same bytes, same taps
    a, ta, _ = emu[4].rx_aligned_batch(rec, taps=True)
    emu[4].use_tma = 0
    b, tb, _ = emu[4].rx_aligned_batch(rec, taps=True)
    emu[4].use_tma = 1
    assert np.array_equal(a, b) and np.array_equal(a, pay)
    assert np.array_equal(ta["constell"], tb["constell"]) and np.array_equal(ta["scal"], tb["scal"])


def test_loopback_clean(emu, port):
    """tx kernel -> rx kernel, noiseless: every payload byte comes back (all modTypes)"""
    for mt in (1, 2, 4, 6, 8):
        s = port[mt].sizes
        pay = pc.synth.payloads(2, s.usefull_size, seed=mt)
        fr = emu[mt].tx_batch(pay, 1)
        out, _ = emu[mt].rx_aligned_batch(fr.reshape(-1, 2), n_frames=2, frame_stride=s.output_size, offset=s.t2sin_size)
        assert np.array_equal(out, pay), mt


def test_sync_kernels_on_reference_capture(emu, port, golden_capture):
    cap16 = golden_capture["capture_i16"][:40960]
    cap = pc.cplx(cap16)
    o = port[4]
    for x in (cap16, cap.astype(np.complex64)):
        rel = emu[4].t2sin_metric(x)
        want = o.t2sin_corr(cap)
        assert np.nonzero(rel > 0.8)[0].tolist() == np.nonzero(want)[0].tolist() == [42, 74]
        assert np.abs(rel[want > 0] - want[want > 0]).max() < 1e-6
        # fewer than 64 blocks take the one-block-per-warp kernel, an odd count leaves the last warp half full
        assert np.abs(emu[4].t2sin_metric(x[:63 * 256]) - rel[:63]).max() < 1e-6
        assert np.abs(emu[4].t2sin_metric(x[:75 * 256]) - rel[:75]).max() < 1e-6
        assert emu[4].find_t2sin(x, 0) == 10752 and emu[4].find_t2sin(x, 11000) == o.find_t2sin(cap, 11000)
        starts = np.array([10752, 18976, 5000, 0])
        first, cor = emu[4].preamble_search(x, starts, want_cor=True)
        assert first.tolist() == [o.find_preamble(cap, int(p)) for p in starts] == [11039, 19301, -10, -10]
        for p, c in zip(starts, cor):
            assert np.abs(c - o.find_corr(cap, int(p))).max() < 1e-6
        emu[4].set_pc_plain(1)                       # the fallback kernel for sizes that are not multiples of 4
        try:
            first2, cor2 = emu[4].preamble_search(x, starts, want_cor=True)
        finally:
            emu[4].set_pc_plain(0)
        assert first2.tolist() == first.tolist() and np.abs(cor2 - cor).max() < 1e-6


def test_rx_phase_unwrap_slow_path(emu, port):
    """timing several samples early => the preamble phase ramps through +-pi more than once, so the
    one-step unwrap of chan_char_lq (Frame.hpp:407-414) takes its slow (chain) path"""
    pay, rec = pc.impaired_records(port[4], 4, seed=77, early=7, noise=0.5)
    st = pc.check_rx_against_oracle(emu[4], port[4], rec, "i16")
    assert st["shift_mismatch"] == 0 and st["differing"] == 0


@pytest.mark.parametrize("mt", [2, 4, 6])
def test_read_syncless_and_chan_char(emu, port, mt):
    st = pc.check_read_and_chan_char(emu[mt], port[mt])
    assert st["rel_l2"] < 2e-6


def test_generic_path_small_geometry(cfg_dir, oracle_lib):
    """the any-size kernels (generic.cuh) on a 128-point configuration: tx and the full rx chain vs the oracle"""
    o = oracle_lib.Oracle("port", cfg_dir["small"])
    m = EmuModem(cfg_dir["small"], o.sizes)
    assert not m.fused
    st = pc.check_tx(m, o, n_frames=2)
    assert st["rel_l2"] < 1e-6
    pay, rec = pc.impaired_records(o, 3, seed=5, cfo_max=0.004, noise=1.0, taps=(1.0, 0.1j), early=1)
    st = pc.check_rx_against_oracle(m, o, rec, "i16")
    assert st["shift_mismatch"] == 0 and st["constell"] < 3e-6
    fr = m.tx_batch(pay, 1)
    s = o.sizes
    out, _ = m.rx_aligned_batch(fr.reshape(-1, 2), n_frames=3, frame_stride=s.output_size, offset=s.t2sin_size)
    assert np.array_equal(out, pay)
    m.close()


def test_production_instantiation_matches_tapped_one(emu, port):
    """no-taps kernel (pruned last FFT pass, ambiguity counting optional) decodes the same bytes"""
    for mt in (2, 4, 6):
        pay, rec = pc.impaired_records(port[mt], 3, seed=31 + mt)
        a, taps, amb_t = emu[mt].rx_aligned_batch(rec, taps=True)
        b, amb = emu[mt].rx_aligned_batch(rec)
        c, _ = emu[mt].rx_aligned_batch(rec, count_ambiguous=False)
        assert np.array_equal(a, b) and np.array_equal(a, c) and amb == amb_t


def test_stream_scan_kernel_matches_reference_loop(cfg_dir, oracle_lib):
    """stream.cuh: rx.cpp's acquisition state machine run by one CTA per shard, against the oracle's loop;
    one shard = the sequential loop, two shards merge to the same list."""
    from cofdm_b200 import stream as st
    o = oracle_lib.Oracle("port", cfg_dir["stream"])
    s = o.sizes
    m = EmuModem(cfg_dir["stream"], s)
    pay = pc.synth.payloads(12, s.usefull_size, seed=3)
    tx16 = np.stack([o.tx(p)[1] for p in pay]).reshape(12, -1, 2)
    rng = np.random.default_rng(11)
    fr = pc.synth.channel(tx16, seed=4, cfo=rng.uniform(-0.003, 0.003, 12), phase=rng.uniform(0, 1, 12), noise_sigma=1.0)
    cap, _ = pc.synth.capture(fr, gaps=rng.integers(260, 9000, 12), noise_sigma=3.0, seed=5, tail=s.output_size * 12)
    want_pos, _ = o.rx_stream(cap)
    assert len(want_pos) >= 9
    blk = st.block_samples(s)
    n_blocks = cap.shape[0] // blk
    assert n_blocks >= 2
    (got,) = m.stream_scan(cap, [(0, n_blocks)])
    assert got.tolist() == want_pos.tolist()
    shards = []
    for r in range(2):
        s0, s1, b0, b1 = st.shard_slice(cap.shape[0], s, r, 2)
        shards.append((s0, (s1 - s0) // blk, b0, b1))
    lists = m.stream_scan(cap, [(a, n) for a, n, _, _ in shards])
    dummy = [np.zeros((len(l), s.usefull_size), np.uint8) for l in lists]
    pos, _, unmerged = st.merge_shards([(l, d, b0, b1) for l, d, (_, _, b0, b1) in zip(lists, dummy, shards)], s)
    assert unmerged == 0 and pos.tolist() == want_pos.tolist()


@pytest.mark.parametrize("ns", [3, 12])
def test_rx_other_symbol_counts(oracle_lib, tmp_path, ns):
    """the demod kernel runs one warp per message symbol: fewer than 8 (part of the shared arrays unused) and more than
    8 (the 15-warp instance) against the oracle, every tap; then the production instance without taps"""
    cfg = pc.synth.write_config(str(tmp_path / f"config_ns{ns}.txt"), num_symb=ns)
    o = oracle_lib.Oracle("port", cfg)
    m = EmuModem(cfg, o.sizes)
    assert m.fused
    pay, rec = pc.impaired_records(o, 2, seed=40 + ns)
    st = pc.check_rx_against_oracle(m, o, rec, "i16")
    assert st["shift_mismatch"] == 0 and st["differing"] == 0
    out, _ = m.rx_aligned_batch(pc.cplx(rec).astype(np.complex64))
    assert np.array_equal(out, np.stack([o.rx_aligned(pc.cplx(r))["bytes"] for r in rec]))
    m.close()


def test_big_path_cluster_kernels(cfg_dir, oracle_lib):
    """BASELINE.json configs[4] (fft 4096, cp 1024, 1920 + 128 sub-carriers, 64-QAM) on the cluster kernels of big.cuh: one
    emulated block of num_symb x 256 threads stands for the cluster of big_demod_kernel (distributed shared memory = offsets);
    acquisition by the any-size kernels (mode 0) and by big_acquire_kernel (mode 1)"""
    o = oracle_lib.Oracle("port", cfg_dir["big"])
    m = EmuModem(cfg_dir["big"], o.sizes)
    assert m.big and not m.fused
    for lay in (1, 0):                                      # big_tx_kernel: the instance compiled for this map and 64-QAM, and the general one
        m.set_big_lay(lay)
        st = pc.check_tx(m, o, n_frames=2)
        assert st["rel_l2"] < 1e-6, (lay, st)
    m.set_big_lay(1)
    pay, rec = pc.impaired_records(o, 2, seed=8, cfo_max=0.0005, noise=0.5, taps=(1.0,), early=0)
    for mode in (0, 1):
        m.big_mode = mode
        st = pc.check_rx_against_oracle(m, o, rec, "i16")
        assert st["shift_mismatch"] == 0 and st["constell"] < 5e-6, (mode, st)
        # the production instances (no taps): compiled for this map's row layout (P.big_lay), and the any-layout one
        assert m.big_lay() == 1
        for lay in (1, 0):
            m.set_big_lay(lay)
            out, _ = m.rx_aligned_batch(pc.cplx(rec).astype(np.complex64), count_ambiguous=False)
            out2, amb = m.rx_aligned_batch(rec, count_ambiguous=True)
            for i in range(len(rec)):
                r = o.rx_aligned(pc.cplx(rec[i]))
                pc.assert_bytes_match(out[i], r["bytes"], r["constell"], o.sizes.mod_type, f"big mode {mode} lay {lay} frame {i}")
                pc.assert_bytes_match(out2[i], r["bytes"], r["constell"], o.sizes.mod_type, f"big mode {mode} lay {lay} i16 frame {i}")
            namb = sum(int(pc.ambiguous_symbols(o.rx_aligned(pc.cplx(x))["constell"], o.sizes.mod_type).sum()) for x in rec)
            assert abs(amb - namb) <= 2, (amb, namb)      # the optional count of decisions near a boundary (margin-level agreement)
        m.set_big_lay(1)


def test_big_path_other_map_and_unaligned_payload(cfg_dir, oracle_lib, tmp_path):
    """an fft-4096 configuration whose sub-carrier map is NOT the one the specialised instances are compiled for (64 pilots,
    960 data sub-carriers: rows 0..2 and 14..15 of the bin matrix) runs the general instances of big_demod_kernel /
    big_tx_kernel; and a payload that is not 16-byte aligned takes big_tx_kernel's byte-load path instead of the bulk copy"""
    cfg = pc.synth.write_config(str(tmp_path / "config_big_np64.txt"), base=cfg_dir["big"], num_data_subc=960, num_pilot_subc=64, num_symb=4)
    o = oracle_lib.Oracle("port", cfg)
    m = EmuModem(cfg, o.sizes)
    assert m.big and m.big_lay() == 0
    st = pc.check_tx(m, o, n_frames=2)
    assert st["rel_l2"] < 1e-6, st
    m.big_mode = 1
    pay, rec = pc.impaired_records(o, 2, seed=12, cfo_max=0.0005, noise=0.5, taps=(1.0,), early=0)
    st = pc.check_rx_against_oracle(m, o, rec, "i16")
    assert st["shift_mismatch"] == 0 and st["constell"] < 5e-6, st
    out, _ = m.rx_aligned_batch(rec, count_ambiguous=False)
    for i in range(len(rec)):
        r = o.rx_aligned(pc.cplx(rec[i]))
        pc.assert_bytes_match(out[i], r["bytes"], r["constell"], o.sizes.mod_type, f"np64 frame {i}")
    m.close()
    # the production map, payload one byte off alignment: same samples as from the aligned copy
    o = oracle_lib.Oracle("port", cfg_dir["big"])
    m = EmuModem(cfg_dir["big"], o.sizes)
    pay = pc.synth.payloads(2, o.sizes.usefull_size, seed=5)
    buf = np.zeros(pay.size + 1, np.uint8)
    buf[1:] = pay.ravel()
    off = buf[1:].reshape(pay.shape)
    assert off.ctypes.data % 16 != 0
    for fmt in (0, 1):
        assert np.array_equal(m.tx_batch(off, fmt), m.tx_batch(pay, fmt))
    m.close()


def test_big_path_phase_unwrap_slow_path(cfg_dir, oracle_lib):
    """a frame cut 6 samples early: the preamble's phases run over several turns, so chan_char_lq's one-step unwrap
    (Frame.hpp:407-414) takes the acquire kernel's 3-state scan path; multipath and a larger CFO on top"""
    o = oracle_lib.Oracle("port", cfg_dir["big"])
    m = EmuModem(cfg_dir["big"], o.sizes)
    m.big_mode = 1
    pay, rec = pc.impaired_records(o, 2, seed=21, cfo_max=0.002, noise=0.5, taps=(1.0, 0.1j), early=6)
    st = pc.check_rx_against_oracle(m, o, rec, "i16")
    assert st["shift_mismatch"] == 0 and st["constell"] < 5e-6 and st["chan"] < 5e-6, st


def test_generic_path_two_preamble_symbols_and_syncless_read(cfg_dir, oracle_lib, tmp_path):
    """configurations the reference accepts and round 1 refused: num_pr_symb = 2 (Frame.cpp:164,259-294: the preamble is two
    OFDM symbols; the coarse spectrum spans both, the phase lock sums over both, the channel fit uses the first) on the
    any-size kernels, and the sync-less FRAME_FORM::read (Frame.cpp:239-242) on the any-size kernels"""
    cfg = pc.synth.write_config(str(tmp_path / "config_small_pr2.txt"), base=cfg_dir["small"], num_pr_symb=2)
    o = oracle_lib.Oracle("port", cfg)
    m = EmuModem(cfg, o.sizes)
    assert not m.fused
    st = pc.check_tx(m, o, n_frames=2)
    assert st["rel_l2"] < 1e-6
    pay, rec = pc.impaired_records(o, 3, seed=31, cfo_max=0.004, noise=1.0, taps=(1.0, 0.1j), early=1)
    st = pc.check_rx_against_oracle(m, o, rec, "i16")
    assert st["shift_mismatch"] == 0 and st["constell"] < 5e-6, st
    # sync-less read on the plain small geometry and on the two-preamble one
    for c in (cfg_dir["small"], cfg):
        o2 = oracle_lib.Oracle("port", c)
        m2 = EmuModem(c, o2.sizes)
        pay2 = pc.synth.payloads(2, o2.sizes.usefull_size, seed=5)
        frames = np.stack([o2.tx(p)[0] for p in pay2]) * 0.8
        out, restored, _, _ = m2.read_batch(frames.astype(np.complex64), taps=True)
        for i in range(2):
            want_b, want_r = o2.read(frames[i])
            assert pc.rel_l2(restored[i], want_r) < 1e-5
            pc.assert_bytes_match(out[i], want_b, want_r, o2.sizes.mod_type, "generic read")
            assert np.array_equal(want_b, pay2[i])


@pytest.mark.parametrize("t2", [128, 512])
def test_t2sin_sizes_other_than_256(oracle_lib, tmp_path, t2):
    """T2sin_size is a configuration key (Frame.cpp:99-136): 128 and 512 run on t2sin_metric_any_kernel; tx writes the tone
    of that size and the detector finds it where the reference does"""
    cfg = pc.synth.write_config(str(tmp_path / f"config_t2_{t2}.txt"), T2sin_size=t2, T2_sin_f1=17 * t2 // 256, T2_sin_f2=51 * t2 // 256)
    o = oracle_lib.Oracle("port", cfg)
    m = EmuModem(cfg, o.sizes)
    s = o.sizes
    assert s.t2sin_size == t2
    st = pc.check_tx(m, o, n_frames=2)
    assert st["rel_l2"] < 1e-6
    pay = pc.synth.payloads(2, s.usefull_size, seed=3)
    tx16 = np.stack([o.tx(p)[1] for p in pay]).reshape(2, -1, 2)
    cap, _ = pc.synth.capture(tx16[..., 0].astype(np.float64) + 1j * tx16[..., 1], gaps=np.array([3 * t2 + 40, 1000]), noise_sigma=2.0, seed=2, tail=4000)
    capc = pc.cplx(cap)
    rel = m.t2sin_metric(cap)
    want = o.t2sin_corr(capc)
    assert np.nonzero(rel > 0.8)[0].tolist() == np.nonzero(want)[0].tolist() and len(np.nonzero(want)[0]) >= 1
    assert np.abs(rel[want > 0] - want[want > 0]).max() < 1e-6
    assert m.find_t2sin(cap, 0) == o.find_t2sin(capc, 0)
