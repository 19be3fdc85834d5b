"""The oracle (oracle/pal_oracle.py) against golden vectors produced by the unmodified
reference (tests/golden/make_golden.py), and the 'restated' layer against the 'port' layer."""
import numpy as np
import pytest

from oracle import pal_oracle as O
from tests.golden.make_golden import CUSTOM_MATERIALS, shoebox


def test_scalars(golden):
    assert O.speed_of_sound(20, 50) == golden["sos_20_50"] == 343.62
    assert O.speed_of_sound(80, 120) == golden["sos_clamped"]
    assert np.array_equal(O.reflect_point_across_plane([0.5, 0.25, 1.5], [1, 2, -1, -3]), golden["reflect"])
    d = O.distance([0.5, 0.5, 0.5], [0, 0, 0])
    assert O.calculate_attenuation(d, "air", 1000, O.DEFAULT_MATERIALS) == golden["att_direct_cfg1"]
    assert golden["att_direct_cfg1"] == 2.8035148612143895e-38
    with pytest.raises(ValueError):
        O.reflect_point_across_plane([0, 0, 0], [0, 0, 0, 1])


def test_image_sources(golden):
    mics, src = golden["img_mics"], golden["img_src"]
    counts = [len(O.generate_image_sources_iterative(src, shoebox(6, 5, 3), o, 1000.0, CUSTOM_MATERIALS,
                                                     mics, absorption_threshold=-1.0)) for o in range(1, 7)]
    assert counts == list(golden["img_counts_thr_neg1"]) == [6, 24, 62, 128, 230, 376]
    im = O.generate_image_sources_iterative(src, shoebox(6, 5, 3), 3, 1000.0, CUSTOM_MATERIALS, mics, 0.01)
    assert np.array_equal(np.array([i["source"] for i in im]), golden["img_pos_o3"])
    names = sorted(CUSTOM_MATERIALS)
    assert [names.index(i["material"]) for i in im] == list(golden["img_mat_o3"])
    cfg1 = O.generate_image_sources_iterative([0.5, 0.5, 0.5],
                                              [{'plane': [1, 0, 0, -5], 'material': 'wood'},
                                               {'plane': [0, 1, 0, -5], 'material': 'metal'},
                                               {'plane': [0, 0, 1, -5], 'material': 'wood'}], 3, 1000,
                                              O.DEFAULT_MATERIALS, np.eye(4, 3, k=-1), 0.01)
    assert len(cfg1) == golden["img_cfg1_count"] == 0


def test_render_port_and_restatement(golden):
    fs = 16000.0
    base = O.generate_signal("chirp", fs, 0.25, 500)
    assert np.array_equal(base, golden["chirp_16k_025_500"])
    assert np.array_equal(O.fractional_delay(np.pad(base, (0, 300)), 0.00731, fs), golden["frac_delay"])
    assert np.array_equal(O.dynamic_range_compression(golden["compress_in"].copy()), golden["compress_out"])
    mics, src = golden["img_mics"][:4], golden["img_src"]
    sig = O.simulate_signals_with_multipath(src, mics, fs, 343.62, duration=0.25, signal_type="chirp", freq=500,
                                            reflective_planes=shoebox(6, 5, 3),
                                            material_properties=CUSTOM_MATERIALS, max_reflections=2,
                                            absorption_threshold=0.01)
    assert np.array_equal(np.array(sig), golden["render_o2_4mics"])
    # restated renderer (shared FFT, transfer-function accumulation, relative gains)
    im = O.generate_image_sources_iterative(src, shoebox(6, 5, 3), 2, 500, CUSTOM_MATERIALS, mics, 0.01)
    tau, gain, total = O.path_table_restated(src, im, mics, fs, 343.62, 0.25, 500, CUSTOM_MATERIALS)
    rows = O.render_rows_restated(base, tau, gain, total, fs, int(0.25 * fs))
    assert np.max(np.abs(rows - golden["render_o2_4mics"])) < 1e-12


def test_gcc_phat_golden(golden):
    r0 = np.random.default_rng(0)
    x = r0.standard_normal(2148)
    s1, s2 = x[100:2148], x[93:2141]
    corr = O.phat_correlation(s1, s2)
    assert np.array_equal(corr, golden["g4_corr"])
    assert int(np.argmax(corr)) == golden["g4_argmax"] == 4088
    for tag, med, lag in (("none", None, 2041), ("0p05", 0.05, 747), ("0p01", 0.01, -103)):
        td, _, _ = O.get_time_delays_phat(s1, s2, 16000, max_expected_delay=med)
        assert np.array_equal(np.array(td), golden[f"g4_td_{tag}"])
        assert td[0] == lag / 16000
        w = O.window_half_width(2048, 2048, 16000, med)
        k = O.tdoa_pick_restated(corr, 2048, w, O.peak_distance(16000))
        assert k[0] - 2047 == lag
        assert O.tdoa_from_index(k[0], 2048, 16000) == td[0]
    td, c0, _ = O.get_time_delays_phat(np.zeros(256), np.zeros(256), 16000, max_expected_delay=0.05)
    assert td[0] == golden["zero_td"][0] == -255 / 16000 and np.max(c0) == golden["zero_max"] == 0
    assert O.tdoa_pick_restated(c0, 256, 800, 16) == [0]


def test_cfg3_frames_golden(golden):
    fr = golden["cfg3_frames"]
    tds, gm = [], []
    for f in range(fr.shape[0]):
        sig = [fr[f, m].astype(np.float64) for m in range(fr.shape[1])]
        t, _, cm = O.pair_loop(sig, 16000.0, max_expected_delay=0.05)
        tds += t
        gm += [cm[i, j] for i in range(8) for j in range(i + 1, 8)]
    assert np.array_equal(np.array(tds), golden["cfg3_td"])
    assert np.array_equal(np.array(gm), golden["cfg3_gmax"])


def test_window_half_width():
    assert O.window_half_width(2048, 2048, 16000, 0.05) == 800
    assert O.window_half_width(44100, 44100, 44100, 0.05) == 2205
    assert O.window_half_width(48000, 48000, 48000, 0.05) == 2400
    assert O.window_half_width(100, 100, 16000, None) == -1
    assert O.window_half_width(100, 100, 16000, -1.0) == -2
    assert O.window_half_width(100, 100, 16000, 1.0) == 100
    with pytest.raises(ValueError):
        O.peak_distance(999)


def test_fuzz_control_flow(golden):
    """Port layer AND restated index algorithm against the reference on 160 fuzzed rows."""
    meta = golden["fuzz_meta"]
    methods = ["median", "adaptive", "other"]
    for i, (n1, n2, fs, mi, mult, med, npk, nout) in enumerate(meta):
        a, b, want = golden[f"fuzz_a{i}"], golden[f"fuzz_b{i}"], golden[f"fuzz_td{i}"]
        med = None if med < 0 else float(med)
        td, corr, _ = O.get_time_delays_phat(a, b, fs, num_peaks=int(npk), threshold_method=methods[int(mi)],
                                             threshold_multiplier=mult, max_expected_delay=med)
        assert np.array_equal(np.array(td), want), i
        w = O.window_half_width(len(a), len(b), fs, med)
        ks = O.tdoa_pick_restated(corr, len(b), w, O.peak_distance(fs), int(npk), methods[int(mi)], mult)
        got = np.array([O.tdoa_from_index(k, len(b), fs) for k in ks])
        assert np.array_equal(got, want), (i, ks)


def test_restated_peak_machinery_vs_scipy():
    from scipy.signal import find_peaks
    rng = np.random.default_rng(3)
    for trial in range(300):
        n = int(rng.integers(3, 300))
        c = rng.standard_normal(n)
        if trial % 3 == 0:
            c = np.round(c * 3) / 3      # plateaus and exact height ties
        dist = int(rng.integers(1, 20))
        pk = O.local_maxima_restated(c)
        assert np.array_equal(pk, find_peaks(c)[0])
        if trial % 3:                    # tie order is only defined for distinct heights
            keep = O.select_by_distance_restated(pk, c[pk], dist)
            assert np.array_equal(pk[keep], find_peaks(c, distance=dist)[0])


def test_synchronize_port_vs_reference_golden(sync_golden):
    """oracle.synchronize_signals_improved (port of utils.py:407-457) against the unmodified reference:
    three scenes incl. an uncorrelated channel (low-peak branch) and fs = 1 kHz."""
    for c in range(int(sync_golden["n_cases"])):
        got = O.synchronize_signals_improved(list(sync_golden[f"in{c}"]), float(sync_golden[f"fs{c}"]))
        assert np.array_equal(np.array(got), sync_golden[f"out{c}"])
