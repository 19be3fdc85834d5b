"""Kernel bodies of pyaudiolocalization_b200/csrc run on the host (tests/emu: the same source
compiled with -DPAL_EMU, OS threads instead of CUDA threads) against the oracle.  This is how
kernel logic is checked in the GPU-less build container; the GPU parity tests proper are in
test_gpu_*.py."""
import numpy as np
import pytest

from oracle import pal_oracle as O
from tests.emu import emu_api as E

N = 4095


def _elem(r, q):
    return (2080 * r + 2016 * q) % N


@pytest.fixture(scope="module")
def frames(golden):
    return np.ascontiguousarray(golden["cfg3_frames"][:1, :5])     # 1 frame, 5 mics (odd: single-channel tail)


@pytest.fixture(scope="module")
def spectra(frames):
    return E.fwd4095(frames)


def test_forward_spectra_layout(frames, spectra):
    """fwd4095 stores the UNIT PHASORS S/|S| of np.fft.fft(sig, 4095) in the [q][r] layout, and per channel
    h = mean_k 1/|S_k|^2, from which the pair kernel bounds the neglected factor |R|/(|R|+1e-10)."""
    for m in range(frames.shape[1]):
        s = np.fft.fft(frames[0, m].astype(np.float64), N)
        want = np.array([[s[_elem(r, q)] for r in range(32)] for q in range(65)]).reshape(-1)
        got = np.asarray(spectra[0, m, :, 0] + 1j * spectra[0, m, :, 1])
        # phase to ~1e-6 rad (relative to the bin's own magnitude the fp32 transform error is |S|max/|S_k| times 1e-7)
        assert np.abs(got * np.abs(want) - want).max() <= 2e-7 * np.abs(want).max() * 8
        assert np.abs(np.abs(got) - 1).max() < 1e-6
        h = np.mean(1.0 / np.abs(s) ** 2)
        assert abs(spectra.hq[0, m, 0] - h) <= 1e-3 * h
        # rounding-noise term of the channel (whiten_bin): mean_k min(4, sigma^2 / |S_k|^2), sigma^2 = 2^-48 sum x^2
        q = np.mean(np.minimum(4.0, 2.0 ** -48 * np.sum(frames[0, m].astype(np.float64) ** 2) / np.abs(s) ** 2))
        assert abs(spectra.hq[0, m, 1] - q) <= 2e-3 * q


def test_whitening_bound_sends_quiet_frames_to_float64():
    """Frames so quiet that |R| is not >> 1e-10 (utils.py:117 is not scale invariant): the whitened fast path
    must flag every row instead of answering, and the bound must dominate the true deviation."""
    rng = np.random.default_rng(5)
    x = rng.standard_normal((1, 3, 2048))
    for scale, expect_flag in ((1.0, False), (1e-2, False), (3e-5, True)):
        fr = (x * scale).astype(np.float32)
        sp = E.fwd4095(fr)
        k, pk, gm, fl, corr = E.pair_fast(sp, E.pairs_of(3), 800, 16, want_corr=True)
        for p, (i, j) in enumerate(E.pairs_of(3)):
            c = O.phat_correlation(fr[0, i].astype(np.float64), fr[0, j].astype(np.float64))
            dev = np.abs(corr[0, p] - c).max()
            bound = 1e-10 * np.sqrt(float(sp.hq[0, i, 0]) * float(sp.hq[0, j, 0]))
            assert dev <= bound + 5e-7
            if expect_flag:
                assert fl[0, p] & 1
            else:
                assert dev < 5e-7


@pytest.mark.parametrize("variant", [0, 2, 3])   # 0: register-resident tiles; 2, 3: tiles parked in (emulated) tensor memory, pair table in shared / global memory
@pytest.mark.parametrize("med", [0.05, 0.01, None])
def test_fast_pair_kernel_vs_oracle(frames, spectra, med, variant):
    pairs = E.pairs_of(frames.shape[1])
    w = O.window_half_width(2048, 2048, 16000.0, med)
    k, pk, gm, fl, corr = E.pair_fast(spectra, pairs, w, 16, want_corr=True, phase_sync=variant)
    for p, (i, j) in enumerate(pairs):
        c = O.phat_correlation(frames[0, i].astype(np.float64), frames[0, j].astype(np.float64))
        assert np.abs(corr[0, p] - c).max() <= 1e-4 * np.abs(c).max()
        assert np.abs(corr[0, p] - c).max() < 5e-7          # what the tie_eps = 2e-6 margin relies on
        want = O.tdoa_pick_restated(c, 2048, w, 16)
        if fl[0, p] & 7 == 0:                                 # unflagged rows must already be exact
            assert k[0, p] == want[0]
        assert abs(gm[0, p] - c.max()) <= 1e-4 * abs(c.max())


@pytest.mark.parametrize("mode", [0, 1, 2])
def test_exact_pair_kernel_vs_oracle(frames, spectra, mode):
    pairs = E.pairs_of(3)
    fr = np.ascontiguousarray(frames[:, :3])
    sp = np.ascontiguousarray(spectra[:, :3])
    k, cnt, pk, gm, fl, corr = E.pair_exact(mode, fr, sp, pairs, 800, 16, num_peaks=3, want_corr=True)
    for p, (i, j) in enumerate(pairs):
        c = O.phat_correlation(fr[0, i].astype(np.float64), fr[0, j].astype(np.float64))
        want = O.tdoa_pick_restated(c, 2048, 800, 16, num_peaks=3)
        assert list(k[0, p, :cnt[0, p]]) == want
        assert np.abs(corr[0, p] - c).max() <= 1e-4 * np.abs(c).max()


def test_exact_kernel_item_list(frames, spectra):
    pairs = E.pairs_of(5)
    k_all, *_ = E.pair_exact(2, frames, spectra, pairs, 800, 16, items=[7, 2])
    for it in (7, 2):
        i, j = pairs[it]
        c = O.phat_correlation(frames[0, i].astype(np.float64), frames[0, j].astype(np.float64))
        assert k_all[0, it, 0] == O.tdoa_pick_restated(c, 2048, 800, 16)[0]
    assert k_all[0, 0, 0] == -7          # untouched rows keep the sentinel


def test_zero_and_identical_channels():
    fr = np.zeros((1, 2, 2048), np.float32)
    sp = E.fwd4095(fr)
    k, pk, gm, fl, _ = E.pair_fast(sp, E.pairs_of(2), 800, 16)
    assert k[0, 0] == 0 and gm[0, 0] == 0 and fl[0, 0] & 16      # SURVEY §8c golden (5): k = 0, max = 0
    k2, cnt, *_ = E.pair_exact(2, fr, sp, E.pairs_of(2), 800, 16)
    assert k2[0, 0, 0] == 0 and cnt[0, 0] == 1


def test_peakpick_fuzz_vs_oracle(golden):
    """The block-level pick (any n, windows, methods, num_peaks, plateaus) on the fuzz rows."""
    meta = golden["fuzz_meta"]
    methods = ["median", "adaptive", "other"]
    checked = 0
    for i in range(0, len(meta), 2):
        n1, n2, fs, mi, mult, med, npk, _ = meta[i]
        a, b = golden[f"fuzz_a{i}"], golden[f"fuzz_b{i}"]
        med = None if med < 0 else float(med)
        c = O.phat_correlation(a, b)
        w = O.window_half_width(len(a), len(b), fs, med)
        d = O.peak_distance(fs)
        want = O.tdoa_pick_restated(c, len(b), w, d, int(npk), methods[int(mi)], mult)
        got, _ = E.peakpick_f64(c, len(b) - 1, w, d, 1 if int(mi) == 1 else 0, mult, int(npk))
        assert got == want, (i, got, want)
        checked += 1
    assert checked >= 80


def test_peakpick_plateaus_and_chains():
    rng = np.random.default_rng(5)
    for trial in range(40):
        n = int(rng.integers(30, 200))
        c = np.round(rng.standard_normal(n) * 2) / 2 if trial % 2 else rng.standard_normal(n)
        if trial % 5 == 0:                       # staircase: long kill chains under the distance rule
            c = np.sort(rng.standard_normal(n)) + 0.5 * ((np.arange(n) % 2) * 2 - 1)
        n2 = n // 2 + 1
        w = int(rng.integers(0, n // 2))
        d = int(rng.integers(1, 12))
        if len(set(np.round(c[O.local_maxima_restated(c)], 12))) < len(O.local_maxima_restated(c)):
            continue                             # equal peak heights: tie order undefined in the reference
        want = O.tdoa_pick_restated(c, n2, w, d, 2)
        got, _ = E.peakpick_f64(c, n2 - 1, w, d, 0, 1.0, 2)
        assert got == want, (trial, got, want)


def test_path_table_batched_matches_per_scene():
    """Batched delay / gain table (one block per scene, in-block max reduction) against the per-scene
    kernel and the oracle restatement of main.py:94-116."""
    from oracle import pal_oracle as O
    from tests.golden.make_golden import CUSTOM_MATERIALS, shoebox
    rng = np.random.default_rng(3)
    mics = rng.uniform([1, 1, 0.5], [5, 4, 2.5], size=(4, 3))
    srcs = rng.uniform([0.5, 0.5, 0.3], [5.5, 4.5, 2.7], size=(3, 3))
    names = list(CUSTOM_MATERIALS)
    mat_abs = [CUSTOM_MATERIALS[n]["absorption"] for n in names]
    mat_freq = [CUSTOM_MATERIALS[n]["freq"] for n in names]
    planes = shoebox(6, 5, 3)
    pl = [p["plane"] for p in planes]
    pm = [names.index(p["material"]) for p in planes]
    pos, mat, cnt = E.image_sources(srcs, pl, pm, mat_abs, mat_freq, mics, 2, 700.0, 0.01, 64)
    tau, gain, pc, mx = E.path_table_batched(srcs, pos, mat, cnt, mics, mat_abs, mat_freq, names.index("air"), 700.0, 343.62)
    for s in range(3):
        k = int(cnt[s])
        assert pc[s] == k + 1
        t1, g1 = E.path_table(srcs[s], pos[s, :k], mat[s, :k], mics, mat_abs, mat_freq, names.index("air"), 700.0, 343.62)
        assert np.array_equal(tau[s, :, :k + 1], t1) and np.array_equal(gain[s, :, :k + 1], g1)
        assert mx[s] == t1.max()
        imgs = O.generate_image_sources_iterative(srcs[s], planes, 2, 700.0, CUSTOM_MATERIALS, mics, 0.01)
        assert len(imgs) == k


@pytest.mark.parametrize("n1,n2,use_double", [(300, 300, False), (257, 190, False), (140, 333, True), (64, 64, True),
                                              (9000, 8200, False), (8300, 9000, True)])     # 256-point tiles: radix-8 steps
def test_generic_bluestein_path_vs_oracle(n1, n2, use_double):
    """Arbitrary-length GCC-PHAT (Bluestein over tiled two-pass FFTs) against the reference algorithm:
    float64 sweep bit-exact on the chosen lag, float32 sweep within 1e-4 on the correlation."""
    rng = np.random.default_rng(n1 * 1000 + n2)
    x = rng.standard_normal(max(n1, n2) + 40)
    a = x[7:7 + n1] + 0.1 * rng.standard_normal(n1)
    b = x[:n2] + 0.1 * rng.standard_normal(n2)
    ld = max(n1, n2)
    sig = np.zeros((1, 2, ld), np.float32)
    sig[0, 0, :n1] = a
    sig[0, 1, :n2] = b
    fs, med = 8000.0, 0.01
    want_td, want_corr, _ = O.get_time_delays_phat(sig[0, 0, :n1].astype(np.float64), sig[0, 1, :n2].astype(np.float64), fs,
                                                   max_expected_delay=med)
    wh = O.window_half_width(n1, n2, fs, med)
    k, _, _, _, _, corr = E.generic_gcc_phat(sig, n1, n2, np.array([[0, 1]], np.int32), wh, O.peak_distance(fs),
                                              use_double=use_double)
    assert np.abs(corr[0, 0] - want_corr).max() <= 1e-4 * np.abs(want_corr).max()
    if use_double:
        assert O.tdoa_from_index(int(k[0, 0, 0]), n2, fs) == want_td[0]


def test_render_scene_emulation_vs_oracle():
    """Renderer kernels (transfer function, Hermitian Bluestein inverse, fade / trim, normalise +
    compress) on the host emulation against the oracle port of main.py:66-124."""
    from tests.golden.make_golden import CUSTOM_MATERIALS, shoebox
    rng = np.random.default_rng(9)
    mics = rng.uniform([1, 1, 0.5], [5, 4, 2.5], size=(3, 3))
    src = rng.uniform([0.5, 0.5, 0.3], [5.5, 4.5, 2.7], size=3)
    fs, dur, freq, c = 8000.0, 0.05, 600.0, 343.62
    planes = shoebox(6, 5, 3)
    imgs = O.generate_image_sources_iterative(src, planes, 1, freq, CUSTOM_MATERIALS, mics, 0.01)
    tau, gain, total = O.path_table_restated(src, imgs, mics, fs, c, dur, freq, CUSTOM_MATERIALS)
    base = O.generate_signal("chirp", fs, dur, freq)
    got = E.render_scene(base, total, tau, gain, fs, int(dur * fs))
    want = np.array(O.simulate_signals_with_multipath(src, mics, fs, c, duration=dur, signal_type="chirp", freq=freq,
                                                      reflective_planes=planes, material_properties=CUSTOM_MATERIALS,
                                                      max_reflections=1, absorption_threshold=0.01))
    assert np.abs(got - want).max() <= 1e-5


@pytest.mark.parametrize("rows,n", [(3, 200), (37, 1000), (1, 2048)])
def test_filtfilt_bit_exact_vs_scipy(rows, n):
    """The channel filter kernel (float64, scipy's direct-form-II-transposed update in its own evaluation
    order) must reproduce scipy.signal.filtfilt bit for bit: the Butterworth band-pass of
    signal_processing.noise_reduction (order 5, 300-3400 Hz) and a low-order filter."""
    from scipy.signal import butter, filtfilt, lfilter_zi
    rng = np.random.default_rng(rows * n)
    x = rng.standard_normal((rows, n))
    for (b, a) in (butter(5, [300 / 8000.0, 3400 / 8000.0], btype="band"), butter(2, 0.2)):
        want = filtfilt(b, a, x, axis=-1)
        got = E.filtfilt_f64(x, b, a, lfilter_zi(b, a), 3 * max(len(a), len(b)))
        assert np.array_equal(got, want)


def _sync_via_emulation(sigs, fs, use_interpolation=True):
    """The product's alignment path with the device kernels emulated on the host."""
    from pyaudiolocalization_b200 import sync as S
    lens = [len(s) for s in sigs]
    n = max(lens)
    host = np.zeros((1, len(sigs), n))
    for i, s in enumerate(sigs):
        host[0, i, :lens[i]] = s
    ref, pk, am, win, en = E.sync_align(host, None if min(lens) == n else np.array([lens], np.int32))
    pads = S.pads_from_alignment(int(ref[0]), pk[0], am[0], win[0], lens, fs, use_interpolation)
    n_out = int(max(p + l for p, l in zip(pads, lens)))
    out = E.pad_rows_f64(host[0], pads, n_out, None if min(lens) == n else np.array(lens, np.int32))
    return out, dict(ref=int(ref[0]), pk=pk[0], am=am[0], win=win[0], en=en[0], pads=pads)


def test_sync_align_emulation_vs_reference_golden(sync_golden):
    """Alignment kernels (energies, float64 Bluestein cross-correlation, arg-max + spline window, padding)
    against the UNMODIFIED reference's synchronize_signals_improved: the aligned channels are bit-identical."""
    from scipy.signal import correlate
    for c in range(int(sync_golden["n_cases"])):
        sigs, fs = list(sync_golden[f"in{c}"]), float(sync_golden[f"fs{c}"])
        out, d = _sync_via_emulation(sigs, fs)
        assert np.array_equal(out, sync_golden[f"out{c}"])
        ref = sigs[d["ref"]]
        assert d["ref"] == int(np.argmax([np.sum(s ** 2) for s in sigs]))
        for m, s in enumerate(sigs):
            corr = correlate(s, ref, mode="full")
            k = int(np.argmax(np.abs(corr)))
            assert d["pk"][m] == k
            assert abs(d["am"][m] - abs(corr[k])) <= 1e-12 * abs(corr[k])
            assert np.abs(d["win"][m] - corr[k - 2:k + 3]).max() <= 1e-11 * abs(corr[k])


def test_sync_align_emulation_unequal_lengths_and_edges():
    """Channels of different lengths (scipy's 'full' index depends on both lengths) and a peak at the edge of
    the row, against the oracle port."""
    rng = np.random.default_rng(11)
    src = rng.standard_normal(700)
    sigs = [src[20:520] * 1.3, src[5:455] + 0.05 * rng.standard_normal(450), src[33:420], 0.01 * rng.standard_normal(300)]
    out, d = _sync_via_emulation(sigs, 4000.0)
    want = O.synchronize_signals_improved([s.copy() for s in sigs], 4000.0)
    assert np.array_equal(out, np.array(want))
    # identical channels (BASELINE cfg1: every microphone hears the same signal): first arg-max, no shift
    same = [src[:256].copy() for _ in range(3)]
    out2, d2 = _sync_via_emulation(same, 16000.0)
    assert d2["ref"] == 0 and np.array_equal(out2, np.array(O.synchronize_signals_improved(same, 16000.0)))


@pytest.mark.parametrize("use_double", [False, True])
def test_generic_path_packs_two_real_sequences_per_transform(use_double):
    """Two channels share a forward transform and two pair-correlations share an inverse one (odd channel count,
    odd item count, several frames): every row must still be the reference's row, with no leakage between the
    two correlation rows of a packed transform."""
    rng = np.random.default_rng(77)
    b, m, n = 3, 3, 180
    sig = np.zeros((b, m, n), np.float32)
    for f in range(b):
        src = rng.standard_normal(n + 30)
        for c in range(m):
            d = int(rng.integers(0, 25))
            sig[f, c] = (1.0 + c) * src[d:d + n] + 0.2 * rng.standard_normal(n)      # very different channel levels
    pairs = E.pairs_of(m)
    fs, med = 8000.0, 0.004
    wh = O.window_half_width(n, n, fs, med)
    k, cnt, pk, gm, fl, corr = E.generic_gcc_phat(sig, n, n, pairs, wh, O.peak_distance(fs), use_double=use_double)
    for f in range(b):
        for p, (i, j) in enumerate(pairs):
            want_td, c, _ = O.get_time_delays_phat(sig[f, i].astype(np.float64), sig[f, j].astype(np.float64), fs,
                                                   max_expected_delay=med)
            assert np.abs(corr[f, p] - c).max() <= (2e-6 if use_double else 1e-4) * np.abs(c).max()
            assert abs(gm[f, p] - c.max()) <= 1e-4 * c.max()
            if use_double:
                assert O.tdoa_from_index(int(k[f, p, 0]), n, fs) == want_td[0]


def test_dead_and_weak_channels_next_to_a_loud_one():
    """Two channels share one complex transform.  A dead (all-zero) microphone next to a live one must give the
    reference's all-zero correlation (k = 0, max = 0), not the whitened rounding residue of its partner; a channel
    60 dB below its partner must keep the float32 accuracy it has on its own (exact power-of-two pre-scaling)."""
    rng = np.random.default_rng(3)
    src = rng.standard_normal(2100)
    fr = np.zeros((1, 4, 2048), np.float32)
    fr[0, 0] = src[7:2055]
    fr[0, 2] = 1e-3 * (src[:2048] + 0.3 * rng.standard_normal(2048))       # weak partner of the dead channel 3 ...
    fr[0, 1] = 0.0                                                         # ... and a dead partner of the loud channel 0
    fr[0, 3] = 1e-3 * src[20:2068]
    sp = E.fwd4095(fr)
    assert not np.asarray(sp[0, 1]).any() and not sp.hq[0, 1].any()
    pairs = E.pairs_of(4)
    k, pk, gm, fl, corr = E.pair_fast(sp, pairs, 800, 16, want_corr=True)
    for p, (i, j) in enumerate(pairs):
        c = O.phat_correlation(fr[0, i].astype(np.float64), fr[0, j].astype(np.float64))
        assert np.abs(corr[0, p] - c).max() < 5e-7, (i, j)
        if 1 in (i, j):
            assert k[0, p] == 0 and gm[0, p] == 0 and fl[0, p] & 16
    # the same through the arbitrary-length (Bluestein) path
    sig = np.ascontiguousarray(fr[:, :, :300])
    kg, cnt, pkg, gmg, flg, cg = E.generic_gcc_phat(sig, 300, 300, pairs, 40, 16)
    for p, (i, j) in enumerate(pairs):
        c = O.phat_correlation(sig[0, i].astype(np.float64), sig[0, j].astype(np.float64))
        assert np.abs(cg[0, p] - c).max() <= 1e-4 * max(np.abs(c).max(), 1e-30) + (0 if c.any() else 0)
        if 1 in (i, j):
            assert not cg[0, p].any() and kg[0, p, 0] == 0


def test_image_sources_cfg4_shape_64_mics_order_6():
    """BASELINE cfg4 geometry: shoebox 6 x 5 x 3 m, 64 microphones, max_reflections = 6 -> 376 images (SURVEY.md 8c
    golden (3)); positions, discovery order and materials bit-identical to the oracle.  64 attenuations per candidate
    go through numpy's pairwise np.mean (eight interleaved accumulators), which the kernel reproduces."""
    from oracle import pal_oracle as O
    from tests.golden.make_golden import CUSTOM_MATERIALS, shoebox
    mics = np.random.default_rng(0).uniform([1, 1, 0.5], [5, 4, 2.5], size=(64, 3))
    srcs = np.random.default_rng(1).uniform([0.5, 0.5, 0.3], [5.5, 4.5, 2.7], size=(2, 3))
    names = list(CUSTOM_MATERIALS)
    planes = shoebox(6, 5, 3)
    for thr, expect in ((0.01, None), (-1.0, 376)):      # cfg4's threshold (prunes some), and no pruning at all
        pos, mat, cnt = E.image_sources(srcs, [p["plane"] for p in planes], [names.index(p["material"]) for p in planes],
                                        [CUSTOM_MATERIALS[n]["absorption"] for n in names],
                                        [CUSTOM_MATERIALS[n]["freq"] for n in names], mics, 6, 1000.0, thr, 400)
        for s in range(2):
            imgs = O.generate_image_sources_iterative(srcs[s], planes, 6, 1000.0, CUSTOM_MATERIALS, mics, thr)
            k = len(imgs)
            assert cnt[s] == k and (expect is None or k == expect) and k > 200
            assert np.array_equal(pos[s, :k], np.array([im["source"] for im in imgs]))
            assert [names[i] for i in mat[s, :k]] == [im["material"] for im in imgs]


def test_numpy_norm_and_mean_rounding_order():
    """The two numpy facts the image-source kernel relies on for its prune test (pal_render.cuh: norm3,
    numpy_pairwise_sum): np.linalg.norm of a 3-vector is sqrt(fma(z, z, fma(y, y, x*x))), and np.mean of 8..128 values
    sums eight interleaved accumulators.  A numpy / BLAS build that rounds differently shows up here, not as a
    silent 1-ulp prune difference."""
    from fractions import Fraction as F
    rng = np.random.default_rng(5)

    def fma(a, b, c):
        return float(F(float(a)) * F(float(b)) + F(float(c)))
    for _ in range(400):
        d = rng.uniform(-10, 10, 3)
        assert float(np.linalg.norm(d)) == float(np.sqrt(fma(d[2], d[2], fma(d[1], d[1], d[0] * d[0]))))
    for n in (5, 8, 13, 64):
        for _ in range(50):
            x = rng.uniform(0, 1, n)
            if n < 8:
                want = 0.0
                for t in x:
                    want += t
            else:
                r = list(x[:8])
                i = 8
                while i < n - n % 8:
                    for j in range(8):
                        r[j] += x[i + j]
                    i += 8
                want = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]))
                for t in x[i:]:
                    want += t
            assert float(np.mean(list(x))) == want / n


@pytest.mark.parametrize("plan,n1,n2", [(0, 700, 650), (1, 1500, 1500), (2, 4000, 4000), (3, 5000, 6100), (4, 8000, 8000),
                                        (5, 9000, 8200), (6, 1200, 1300), (7, 44100, 44100), (8, 3000, 2500)])
def test_fft2_engine_every_plan_vs_oracle(plan, n1, n2):
    """Second-generation float32 convolution engine (pal_fft2.cuh: compile-time plans M1 x M2 incl. 3 * 2^k lengths,
    radix-3/8/16 register steps, fused row pass) through the whole GCC-PHAT path against the reference algorithm, one
    case per plan: correlation within 1e-4 of max|corr| (float32) and the reference's lag.  Plan 7 (384 x 512 = 196608)
    at n = 88199 is BASELINE cfg2's transform."""
    rng = np.random.default_rng(n1 + plan)
    x = rng.standard_normal(max(n1, n2) + 40)
    ld = max(n1, n2)
    sig = np.zeros((1, 2, ld), np.float32)
    sig[0, 0, :n1] = x[7:7 + n1] + 0.1 * rng.standard_normal(n1)
    sig[0, 1, :n2] = x[:n2] + 0.1 * rng.standard_normal(n2)
    fs, med = 8000.0, 0.01
    want_td, want_corr, _ = O.get_time_delays_phat(sig[0, 0, :n1].astype(np.float64), sig[0, 1, :n2].astype(np.float64), fs,
                                                   max_expected_delay=med)
    k, _, _, _, _, corr, used = E.fft2_gcc_phat(sig, n1, n2, np.array([[0, 1]], np.int32), O.window_half_width(n1, n2, fs, med),
                                                 O.peak_distance(fs), plan_id=plan)
    assert used == plan
    assert np.abs(corr[0, 0] - want_corr).max() <= 1e-4 * np.abs(want_corr).max()
    assert O.tdoa_from_index(int(k[0, 0, 0]), n2, fs) == want_td[0]


def test_fft2_packs_channels_and_pairs_like_the_first_engine():
    """4 channels / 6 pairs of 2 frames (two real sequences per complex transform in both directions, a dead channel in
    one frame): the second-generation engine must agree with the first-generation float32 sweep to float32 rounding and
    pick the same lags."""
    rng = np.random.default_rng(12)
    n = 1600
    src = rng.standard_normal((2, n + 32))
    sig = np.stack([np.stack([src[f, 8 * c:8 * c + n] + 0.2 * rng.standard_normal(n) for c in range(4)]) for f in range(2)]).astype(np.float32)
    sig[1, 2] = 0.0
    pairs = E.pairs_of(4)
    a = E.generic_gcc_phat(sig, n, n, pairs, 80, 8)
    b = E.fft2_gcc_phat(sig, n, n, pairs, 80, 8)
    assert np.array_equal(a[0], b[0])
    assert np.abs(a[5] - b[5]).max() <= 2e-6
    assert not b[5][1, [1, 3, 5]].any()          # pairs with the dead channel: exact zeros


@pytest.mark.parametrize("med", [0.02, 0.003, None])
def test_reduced_window_pick_vs_oracle(med):
    """pal_winpick.cuh: the inverse column pass keeps only the window (+ `distance` margin) and per-tile row maxima; a
    warp picks the window maximum and flags everything a float64 look could change.  Every UNFLAGGED row must already
    be the reference's answer, max(corr) must match on every row, and few rows may be flagged."""
    rng = np.random.default_rng(31)
    b, m, n, fs = 12, 4, 1400, 16000.0
    src = rng.standard_normal((b, n + 64))
    d = rng.integers(0, 40, size=(b, m))
    sig = np.stack([np.stack([src[f, 40 - d[f, c]:40 - d[f, c] + n] for c in range(m)]) for f in range(b)])
    sig = (sig + 0.3 * rng.standard_normal(sig.shape)).astype(np.float32)
    sig[3, 1] = 0.0                                                   # a dead channel: rows of exact zeros
    pairs = E.pairs_of(m)
    wh = O.window_half_width(n, n, fs, med)
    k, cnt, pk, gm, fl, _, _ = E.fft2_gcc_phat(sig, n, n, pairs, wh, O.peak_distance(fs), eps=1e-6, fast=True)
    flagged = (fl & 1) != 0
    for f in range(b):
        for p, (i, j) in enumerate(pairs):
            td, corr, _ = O.get_time_delays_phat(sig[f, i].astype(np.float64), sig[f, j].astype(np.float64), fs, max_expected_delay=med)
            assert abs(gm[f, p] - corr.max()) <= 1e-4 * max(abs(corr.max()), 1e-30)
            if not flagged[f, p]:
                assert O.tdoa_from_index(int(k[f, p, 0]), n, fs) == td[0], (f, p)
                assert abs(pk[f, p] - corr[k[f, p, 0]]) <= 1e-4 * abs(corr).max()
    assert flagged[3, [0, 3, 4]].all()                                # zero rows: the reference's argmax fallback decides
    # (a 97-sample window holds only noise-level samples -- the reference's window is centred on IFFT index n2-1, not
    # on lag 0 -- and its maximum is not always clear of the mean|c| bound: those rows go to the float64 sweep)
    assert flagged.mean() < (0.35 if med == 0.003 else 0.12)


def test_batched_position_solve_vs_scipy():
    """pal_solver.cuh on the host emulation: the reference's residuals (utils.py:384-405) and dynamic box
    (utils.py:364-382) solved per scene by one warp; positions within 1e-6 m of scipy's bounded trf run to tight
    tolerances from the same start."""
    from scipy.optimize import least_squares
    rng = np.random.default_rng(17)
    s_n, m, c = 12, 6, 343.62
    dims = rng.uniform([3, 3, 2.5], [10, 8, 4], size=(s_n, 3))
    mics = 0.3 + rng.uniform(size=(s_n, m, 3)) * (dims[:, None, :] - 0.6)
    src = 0.3 + rng.uniform(size=(s_n, 3)) * (dims - 0.6)
    pairs = E.pairs_of(m)
    d = np.linalg.norm(src[:, None, :] - mics, axis=2)
    td = (d[:, pairs[:, 1]] - d[:, pairs[:, 0]]) / c + 2e-6 * rng.standard_normal((s_n, len(pairs)))
    w = rng.uniform(0.5, 1.5, size=len(pairs))
    pos, cost, it = E.solve_positions(mics, pairs, td, c, weights=w)

    def eq(x, mc, t):
        dd = np.linalg.norm(x[None, :] - mc, axis=1)
        return ((dd[pairs[:, 1]] - dd[pairs[:, 0]]) - c * t) * w
    same = 0
    for s in range(s_n):
        margin = 5.0 + max(np.percentile(c * np.abs(td[s]), 75), 1.0)
        lo, hi = mics[s].min(axis=0) - margin, mics[s].max(axis=0) + margin
        ref = least_squares(eq, np.clip(mics[s].mean(axis=0), lo, hi), args=(mics[s], td[s]), bounds=(lo, hi), method="trf",
                            ftol=1e-14, xtol=1e-14, gtol=1e-14, max_nfev=2000)
        assert it[s] > 0
        if abs(ref.cost - cost[s]) <= 1e-9 * ref.cost + 1e-18:
            same += 1
            assert np.abs(ref.x - pos[s]).max() <= 1e-6, (s, ref.x, pos[s])
        else:
            assert cost[s] <= ref.cost * (1 + 1e-9)
    assert same >= s_n - 2


def test_batched_position_solve_on_an_active_bound():
    """Sources above the box's ceiling: the minimiser lies on the bound; the active-set step must reach the constrained
    minimum scipy's trf finds (cost equal to 1e-9), not the clipped unconstrained step."""
    from scipy.optimize import least_squares
    rng = np.random.default_rng(3)
    mics = rng.uniform([0, 0, 0], [4, 3, 2], size=(6, 3))
    pairs = E.pairs_of(6)
    srcs = rng.uniform([0.5, 0.5, 0.3], [3.5, 2.5, 1.7], size=(5, 3))
    c = 343.0
    d = np.linalg.norm(srcs[:, None, :] - mics[None], axis=2)
    td = (d[:, pairs[:, 1]] - d[:, pairs[:, 0]]) / c
    lo, hi = np.tile([0.0, 0.0, 0.0], (5, 1)), np.tile([4.0, 3.0, 1.0], (5, 1))
    x0 = np.tile([2.0, 1.5, 0.5], (5, 1))
    pos, cost, it = E.solve_positions(mics, pairs, td, c, x0=x0, lo=lo, hi=hi)

    def eq(x, t):
        dd = np.linalg.norm(x[None, :] - mics, axis=1)
        return (dd[pairs[:, 1]] - dd[pairs[:, 0]]) - c * t
    assert (pos >= lo - 1e-12).all() and (pos <= hi + 1e-12).all()
    for s in range(5):
        ref = least_squares(eq, x0[s], args=(td[s],), bounds=(lo[s], hi[s]), method="trf", ftol=1e-14, xtol=1e-14, gtol=1e-14)
        assert cost[s] <= ref.cost + 1e-9, (s, cost[s], ref.cost)
        if srcs[s, 2] <= 1.0:
            assert np.abs(pos[s] - srcs[s]).max() < 1e-6


def test_fast_path_whitening_bound_flags_quiet_channels():
    """Per-channel whitening of the fast path (pal_winpick.cuh): a frame scaled down to 1e-7 of full scale has
    |S_i||S_j| ~ 1e-10, where the reference's absolute 1e-10 in R / (|R| + 1e-10) matters: the whitening bound must
    exceed the near-tie margin and every row of that frame must be handed to the float64 sweep, while the same frame
    at normal level is decided by the fast path (and matches the oracle, checked in the test above)."""
    rng = np.random.default_rng(8)
    n, m, fs = 1400, 3, 16000.0
    src = rng.standard_normal(n + 32)
    fr = np.stack([src[4 * c:4 * c + n] + 0.3 * rng.standard_normal(n) for c in range(m)]).astype(np.float32)
    sig = np.stack([fr, fr * np.float32(1e-7)])
    pairs = E.pairs_of(m)
    k, cnt, pk, gm, fl, _, _ = E.fft2_gcc_phat(sig, n, n, pairs, O.window_half_width(n, n, fs, 0.02), O.peak_distance(fs),
                                               eps=1e-6, fast=True)
    assert ((fl[1] & 1) != 0).all()               # quiet frame: all rows flagged for the float64 sweep
    assert ((fl[0] & 1) == 0).sum() >= 2          # normal level: decided by the fast path


@pytest.mark.parametrize("plan,n1,n2", [(0, 700, 650), (1, 1500, 1500), (2, 4000, 4000)])
def test_single_cta_convolution_vs_oracle(plan, n1, n2):
    """pal_fft2.cuh conv_smem_body: the whole convolution (column FFT, twiddle, row FFT, chirp spectrum, inverse) in one
    CTA's shared memory, for the plans of at most 16384 points; through the whole GCC-PHAT path against the reference
    algorithm, and equal to the three-kernel engine to float32 rounding."""
    rng = np.random.default_rng(n1 + plan)
    x = rng.standard_normal(max(n1, n2) + 40)
    ld = max(n1, n2)
    sig = np.zeros((2, 2, ld), np.float32)
    for f in range(2):
        sig[f, 0, :n1] = x[7 + f:7 + f + n1] + 0.1 * rng.standard_normal(n1)
        sig[f, 1, :n2] = x[:n2] + 0.1 * rng.standard_normal(n2)
    fs, med = 8000.0, 0.01
    pairs = np.array([[0, 1]], np.int32)
    wh, dist = O.window_half_width(n1, n2, fs, med), O.peak_distance(fs)
    a = E.fft2_gcc_phat(sig, n1, n2, pairs, wh, dist, plan_id=plan, smem_conv=True)
    b = E.fft2_gcc_phat(sig, n1, n2, pairs, wh, dist, plan_id=plan)
    assert np.abs(a[5] - b[5]).max() <= 2e-6 and np.array_equal(a[0], b[0])
    for f in range(2):
        want_td, want_corr, _ = O.get_time_delays_phat(sig[f, 0, :n1].astype(np.float64), sig[f, 1, :n2].astype(np.float64), fs,
                                                       max_expected_delay=med)
        assert np.abs(a[5][f, 0] - want_corr).max() <= 1e-4 * np.abs(want_corr).max()
        assert O.tdoa_from_index(int(a[0][f, 0, 0]), n2, fs) == want_td[0]


def test_rounding_noise_bound_covers_high_dynamic_range_frames():
    """Windowed tones over a -80 dB floor: the float32 transform's rounding noise turns the phases of the stop-band bins
    (which PHAT weights like any other) by ~1e-3 rad, and the correlation moves by far more than the tie margin.  The
    per-channel term q (whiten_bin) must bound the deviation row by row, send such rows to the float64 kernel, and
    leave ordinary frames alone."""
    rng = np.random.default_rng(12)
    n, fs = 2048, 16000.0
    t = np.arange(n) / fs
    tones = np.zeros((1, 3, n), np.float32)
    for c in range(3):
        d = int(rng.integers(0, 300))
        x = sum(np.sin(2 * np.pi * f * (t - d / fs)) for f in (440.0, 1234.5, 0.21 * fs))
        tones[0, c] = x * np.hanning(n) + 1e-4 * rng.standard_normal(n)
    noise = rng.standard_normal((1, 3, n)).astype(np.float32)
    pairs = E.pairs_of(3)
    for fr, high in ((tones, True), (noise, False)):
        sp = E.fwd4095(fr)
        k, pk, gm, fl, corr = E.pair_fast(sp, pairs, 800, 16, eps=1e-6, want_corr=True)
        for p, (i, j) in enumerate(pairs):
            c = O.phat_correlation(fr[0, i].astype(np.float64), fr[0, j].astype(np.float64))
            dev = np.abs(corr[0, p] - c).max()
            bound = 20.0 * np.sqrt((float(sp.hq[0, i, 1]) + float(sp.hq[0, j, 1])) / N)
            assert dev <= bound / 2.5 + 2e-7, (high, i, j, dev, bound)
            if high:
                assert dev > 1e-5 and bound > 1e-4 * c.max() and (fl[0, p] & 1)
            else:
                assert bound < 2e-7
