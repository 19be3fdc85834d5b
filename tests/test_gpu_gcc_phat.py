"""GPU parity tests of the GCC-PHAT / TDOA path (run with -m gpu on the B200 box).
Everything goes through the C ABI (pal_gcc_phat_tdoa via pyaudiolocalization_b200.gcc_phat).
Bars: integer lag indices and TDOAs bit-exact vs the reference algorithm; correlation values
within 1e-4 relative (to max|corr| of the row), fp32."""
import numpy as np
import pytest

from oracle import pal_oracle as O

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

CORR_RTOL = 1e-4


@pytest.fixture(scope="module")
def pal():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import pyaudiolocalization_b200 as p
    return p


def _oracle_rows(fr, fs, med, **kw):
    """reference answer for every pair of every frame: (td list, corr max, k list)"""
    b, m, n = fr.shape
    out = []
    for f in range(b):
        for i in range(m):
            for j in range(i + 1, m):
                td, corr, _ = O.get_time_delays_phat(fr[f, i].astype(np.float64), fr[f, j].astype(np.float64),
                                                     fs, max_expected_delay=med, **kw)
                out.append((td, corr))
    return out


def test_golden_cfg3_frames(pal, golden):
    """TDOAs of the committed cfg3-shaped frames must equal what the unmodified reference printed."""
    fr = golden["cfg3_frames"]
    res = pal.gcc_phat_tdoa_batched(torch.from_numpy(fr).cuda(), 16000.0, max_expected_delay=0.05)
    td = res.tdoa_seconds()[..., 0].reshape(-1)
    assert np.array_equal(td, golden["cfg3_td"])
    gm = res.gmax.cpu().numpy().reshape(-1)
    assert np.max(np.abs(gm - golden["cfg3_gmax"]) / np.abs(golden["cfg3_gmax"])) < CORR_RTOL
    assert int((res.k_count != 1).sum()) == 0


@pytest.mark.parametrize("med", [None, 0.05, 0.01, 0.0005, 1.0])
def test_random_frames_vs_oracle(pal, med):
    fr = pal.synth.cfg3_frames(3, mics=6, seed=11).cpu().numpy() if hasattr(pal, "synth") else None
    if fr is None:
        from pyaudiolocalization_b200 import synth
        fr = synth.cfg3_frames(3, mics=6, seed=11).cpu().numpy()
    res = pal.gcc_phat_tdoa_batched(torch.from_numpy(fr).cuda(), 16000.0, max_expected_delay=med,
                                    return_corr=True)
    td = res.tdoa_seconds()[..., 0].reshape(-1)
    corr = res.corr.cpu().numpy().reshape(-1, 4095)
    rows = _oracle_rows(fr, 16000.0, med)
    for r, (want_td, want_corr) in enumerate(rows):
        assert td[r] == want_td[0], (r, td[r], want_td)
        assert np.abs(corr[r] - want_corr).max() <= CORR_RTOL * np.abs(want_corr).max()


@pytest.mark.parametrize("kw", [dict(num_peaks=3), dict(threshold_method="adaptive", threshold_multiplier=3.0),
                                dict(threshold_multiplier=8.0), dict(num_peaks=2, threshold_method="adaptive")])
def test_options_vs_oracle(pal, kw):
    from pyaudiolocalization_b200 import synth
    fr = synth.cfg3_frames(2, mics=4, seed=5).cpu().numpy()
    res = pal.gcc_phat_tdoa_batched(torch.from_numpy(fr).cuda(), 16000.0, max_expected_delay=0.02, **kw)
    td = res.tdoa_seconds().reshape(-1, kw.get("num_peaks", 1))
    cnt = res.k_count.cpu().numpy().reshape(-1)
    rows = _oracle_rows(fr, 16000.0, 0.02, **kw)
    for r, (want_td, _) in enumerate(rows):
        assert cnt[r] == len(want_td)
        assert np.array_equal(td[r, :cnt[r]], np.array(want_td)), (r, td[r], want_td)


def test_degenerate_rows(pal):
    fr = np.zeros((1, 3, 2048), np.float32)
    fr[0, 2] = np.random.default_rng(0).standard_normal(2048)
    res = pal.gcc_phat_tdoa_batched(torch.from_numpy(fr).cuda(), 16000.0, max_expected_delay=0.05)
    k = res.k_idx.cpu().numpy()[0, :, 0]
    gm = res.gmax.cpu().numpy()[0]
    assert k[0] == 0 and gm[0] == 0.0          # zeros x zeros: SURVEY §8c golden (5)
    assert k[1] == 0 and gm[1] == 0.0          # zeros x noise: R == 0 as well
    with pytest.raises(ValueError):
        pal.gcc_phat_tdoa_batched(torch.from_numpy(fr).cuda(), 999.0)     # scipy: distance < 1


def test_chunked_equals_single_pass_and_flag_budget(pal):
    """Size-independent properties at a size the oracle cannot cover: determinism, chunking
    invariance, flag rates, and every unflagged in-window answer is a strict local maximum
    that dominates its window."""
    from pyaudiolocalization_b200 import synth
    fr = synth.cfg3_frames(256, mics=32, seed=3000)
    a = pal.gcc_phat_tdoa_batched(fr, 16000.0, max_expected_delay=0.05, return_corr=False)
    full, small = pal.gcc_phat.workspace_bytes(256, 32, 2048, 496)
    b = pal.gcc_phat_tdoa_batched(fr, 16000.0, max_expected_delay=0.05, max_workspace_bytes=small + 40 * 532480)
    assert torch.equal(a.k_idx, b.k_idx) and torch.equal(a.peak, b.peak) and torch.equal(a.flags, b.flags)
    c = pal.gcc_phat_tdoa_batched(fr, 16000.0, max_expected_delay=0.05)
    assert torch.equal(a.k_idx, c.k_idx)
    flags = a.flags.cpu().numpy()
    refined = (flags & 8) != 0
    assert refined.mean() < 0.01, refined.mean()
    lag = a.lags().cpu().numpy()[..., 0]
    inwin = np.abs(lag) <= 800
    assert (inwin | ((flags & 16) != 0)).all()          # outside the window only via the argmax fallback
    # refine on/off may only differ on flagged rows
    d = pal.gcc_phat_tdoa_batched(fr, 16000.0, max_expected_delay=0.05, refine=False)
    diff = (d.k_idx != a.k_idx).cpu().numpy()[..., 0]
    assert not (diff & ~refined).any()
    # a sample of rows against the oracle, flagged rows first
    frh = fr.cpu().numpy()
    pairs = pal.all_pairs(32)
    idx = list(np.argwhere(refined)[:12]) + [np.array([f, p]) for f, p in zip(range(0, 256, 16), range(0, 496, 31))]
    k = a.k_idx.cpu().numpy()[..., 0]
    for f, p in idx:
        i, j = pairs[p]
        corr = O.phat_correlation(frh[f, i].astype(np.float64), frh[f, j].astype(np.float64))
        want = O.tdoa_pick_restated(corr, 2048, 800, 16)
        assert k[f, p] == want[0], (f, p, k[f, p], want, flags[f, p])


# ------------------------------------------------------------------ arbitrary lengths (Bluestein path)
@pytest.mark.parametrize("n1,n2,m", [(50, 37, 2), (300, 300, 3), (1000, 1000, 2), (4000, 4000, 4), (44100, 44100, 2)])
def test_generic_lengths_vs_oracle(pal, n1, n2, m):
    rng = np.random.default_rng(n1 + n2)
    ld = max(n1, n2)
    fr = np.zeros((2, m, ld), np.float32)
    base = rng.standard_normal(ld + 64)
    for f in range(2):
        for c in range(m):
            ln = n1 if c % 2 == 0 else n2
            d = int(rng.integers(0, 12))
            fr[f, c, :ln] = base[d:d + ln] + 0.3 * rng.standard_normal(ln)
    fs = 16000.0 if ld < 10000 else 44100.0
    med = 0.004
    res = pal.gcc_phat_tdoa_batched(torch.from_numpy(fr).cuda(), fs, max_expected_delay=med, return_corr=True,
                                    lengths=(n1, n2) if n1 != n2 else None)
    td = res.tdoa_seconds()[..., 0]
    corr = res.corr.cpu().numpy()
    pairs = pal.all_pairs(m)
    worst = 0.0
    for f in range(2):
        for p, (i, j) in enumerate(pairs):
            a = fr[f, i, :(n1 if i % 2 == 0 else n2)].astype(np.float64)
            b = fr[f, j, :(n1 if j % 2 == 0 else n2)].astype(np.float64)
            want_td, want_corr, _ = O.get_time_delays_phat(a, b, fs, max_expected_delay=med)
            err = np.abs(corr[f, p] - want_corr).max() / np.abs(want_corr).max()
            worst = max(worst, err)
            assert err <= CORR_RTOL, (f, p, err)
            assert td[f, p] == want_td[0], (f, p, td[f, p], want_td)
    print(f"n1={n1} n2={n2}: worst corr error {worst:.2e} relative to max|corr|")


def test_reference_signature_functions_on_fuzz_rows(pal, golden):
    """utils.get_time_delays_phat / utils.phat_correlation (reference signatures) reproduce the
    TDOA lists the unmodified reference produced for the committed fuzz rows."""
    from pyaudiolocalization_b200 import utils as U
    meta = golden["fuzz_meta"]
    methods = ["median", "adaptive", "other"]
    n_exact = 0
    for i in range(0, len(meta), 3):
        n1, n2, fs, mi, mult, med, npk, _ = meta[i]
        a, b, want = golden[f"fuzz_a{i}"], golden[f"fuzz_b{i}"], golden[f"fuzz_td{i}"]
        med = None if med < 0 else float(med)
        td, corr, lags = U.get_time_delays_phat(a, b, fs, num_peaks=int(npk), threshold_method=methods[int(mi)],
                                                threshold_multiplier=mult, max_expected_delay=med)
        ref_corr = O.phat_correlation(a, b)
        assert corr.dtype == np.float64 and corr.shape == ref_corr.shape and lags.shape == ref_corr.shape
        scale = max(np.abs(ref_corr).max(), 1e-30)
        assert np.abs(corr - ref_corr).max() <= CORR_RTOL * scale
        assert isinstance(td, list) and np.array_equal(np.array(td), want), (i, td, want)
        n_exact += 1
    assert n_exact >= 50
    c = U.phat_correlation(golden["fuzz_a1"], golden["fuzz_b1"])
    r = O.phat_correlation(golden["fuzz_a1"], golden["fuzz_b1"])
    assert np.abs(c - r).max() <= CORR_RTOL * np.abs(r).max()


def test_host_streaming_entry_point_matches_device_path(pal):
    """gcc_phat_tdoa_from_host (pinned host frames -> chunked H2D / kernels / D2H pipeline) must return
    exactly what the device-resident call returns; the float64 TDOA seconds derived on the device
    (pal_tdoa_seconds) must be bit-identical to numpy's int64 / float64 (utils.py:141-142)."""
    from pyaudiolocalization_b200 import synth
    from pyaudiolocalization_b200.gcc_phat import gcc_phat_tdoa_from_host
    fr = synth.cfg3_frames(37, mics=5, seed=21)
    ref = pal.gcc_phat_tdoa_batched(fr, 16000.0, 0.05)
    host = torch.empty(fr.shape, dtype=torch.float32, pin_memory=True)
    host.copy_(fr)
    torch.cuda.synchronize()
    for chunk in (8, 37, 64):
        r = gcc_phat_tdoa_from_host(host, 16000.0, 0.05, chunk_frames=chunk)
        assert np.array_equal(r["k_idx"], ref.k_idx.cpu().numpy())
        assert np.array_equal(r["gmax"], ref.gmax.cpu().numpy())
        assert r["tdoa"].dtype == np.float64 and np.array_equal(r["tdoa"], ref.tdoa_seconds())
        assert r["h2d_bytes"] == fr.numel() * 4


@pytest.mark.parametrize("n", [2048, 1500])
def test_dead_weak_and_quiet_channels(pal, n):
    """Edge cases of sharing one complex transform between two real channels (both the fused n = 4095 kernels and the
    Bluestein path do): a dead microphone next to a live one gives the reference's all-zero row (k = 0, max = 0), a
    channel 60 dB below its partner keeps its accuracy, and frames so quiet that the absolute 1e-10 of utils.py:117
    matters are still answered exactly (the whitened fast path hands them to the float64 kernel)."""
    rng = np.random.default_rng(3)
    src = rng.standard_normal(n + 100)
    fr = np.zeros((3, 4, n), np.float32)
    fr[0, 0] = src[7:7 + n]
    fr[0, 2] = 1e-3 * (src[:n] + 0.3 * rng.standard_normal(n))
    fr[0, 3] = 1e-3 * src[20:20 + n]                      # channel 1 stays dead
    for c in range(4):                                    # a quiet frame (-90 dBFS) and an ordinary one
        fr[1, c] = 3e-5 * (src[3 * c:3 * c + n] + 0.2 * rng.standard_normal(n))
        fr[2, c] = src[5 * c:5 * c + n] + 0.2 * rng.standard_normal(n)
    res = pal.gcc_phat_tdoa_batched(torch.from_numpy(fr).cuda(), 16000.0, max_expected_delay=0.05, return_corr=True)
    td = res.tdoa_seconds()[..., 0]
    corr, gm, k = res.corr.cpu().numpy(), res.gmax.cpu().numpy(), res.k_idx.cpu().numpy()[..., 0]
    pairs = pal.all_pairs(4)
    for f in range(3):
        for p, (i, j) in enumerate(pairs):
            want_td, c, _ = O.get_time_delays_phat(fr[f, i].astype(np.float64), fr[f, j].astype(np.float64), 16000.0,
                                                   max_expected_delay=0.05)
            assert td[f, p] == want_td[0], (f, i, j)
            if f == 0 and 1 in (i, j):
                assert k[f, p] == 0 and gm[f, p] == 0 and not corr[f, p].any()
            else:
                assert abs(gm[f, p] - c.max()) <= CORR_RTOL * c.max()
                if f != 1:     # (a quiet row's fp32 correlation VALUES deviate by the documented bound; its lag and max are float64)
                    assert np.abs(corr[f, p] - c).max() <= CORR_RTOL * np.abs(c).max()


def test_pcm16_host_frames(pal):
    """int16 PCM host buffers: pal_pcm16_to_f32 is exact (x / 32768, any alignment), and the streaming host entry
    point fed with int16 gives the reference's TDOAs for the float signal soundfile would have produced."""
    from pyaudiolocalization_b200 import gcc_phat as G
    rng = np.random.default_rng(21)
    raw = rng.integers(-32768, 32768, size=100003, dtype=np.int16)
    dev = torch.from_numpy(raw).cuda()
    for off, cnt in ((0, 100003), (1, 99999), (3, 17), (5, 4096), (7, 0)):
        got = G.pcm16_to_f32(dev[off:off + cnt].contiguous() if off == 0 else dev[off:off + cnt]).cpu().numpy()
        assert np.array_equal(got, raw[off:off + cnt].astype(np.float32) / np.float32(32768.0))
    src = rng.standard_normal(2200)
    fr = np.stack([np.stack([src[3 * c + f:3 * c + f + 2048] + 0.2 * rng.standard_normal(2048) for c in range(4)])
                   for f in range(5)])
    pcm = np.clip(np.round(fr * 6000.0), -32768, 32767).astype(np.int16)
    res = G.gcc_phat_tdoa_from_host(torch.from_numpy(pcm).pin_memory(), 16000.0, 0.05, chunk_frames=2)
    assert res["h2d_bytes"] == pcm.size * 2
    x = pcm.astype(np.float64) / 32768.0
    r = 0
    for f in range(5):
        for i in range(4):
            for j in range(i + 1, 4):
                want_td, c, _ = O.get_time_delays_phat(x[f, i], x[f, j], 16000.0, max_expected_delay=0.05)
                assert res["tdoa"][f, r % 6, 0] == want_td[0]
                assert abs(res["gmax"][f, r % 6] - c.max()) <= CORR_RTOL * c.max()
                r += 1


def test_cfg3_parity_subset_every_pair(pal):
    """SURVEY.md section 8d parity subset at the benchmark shape: every one of the 496 pairs of the first 8 frames of the
    cfg3 workload (3968 rows, the generator and seed bench.py uses) against the reference algorithm: TDOA bit-exact,
    max(corr) within 1e-4."""
    from pyaudiolocalization_b200 import synth
    fr = synth.cfg3_frames(8, mics=32, seed=3000)
    res = pal.gcc_phat_tdoa_batched(fr, 16000.0, max_expected_delay=0.05)
    td = res.tdoa_seconds()[..., 0]
    gm = res.gmax.cpu().numpy()
    frh = fr.cpu().numpy().astype(np.float64)
    pairs = pal.all_pairs(32)
    bad = 0
    for f in range(8):
        for p, (i, j) in enumerate(pairs):
            want_td, c, _ = O.get_time_delays_phat(frh[f, i], frh[f, j], 16000.0, max_expected_delay=0.05)
            bad += int(td[f, p] != want_td[0])
            assert abs(gm[f, p] - c.max()) <= CORR_RTOL * c.max()
    assert bad == 0


def test_generic_path_parity_sweep_and_flag_rate(pal):
    """Arbitrary-length path at volume: every pair of 128 frames x 8 mics x 1000 samples (3584 rows, n = 1999) must
    give the reference's TDOA bit for bit, while only a small fraction of the rows may need the float64 sweep (the
    near-tie audit looks at the window and at the winner's `dist`-neighbourhood, not at everything near the window)."""
    rng = np.random.default_rng(99)
    b, m, n, fs, med = 128, 8, 1000, 16000.0, 0.02
    src = rng.standard_normal((b, n + 64)).astype(np.float32)
    d = rng.integers(0, 48, size=(b, m))
    fr = np.stack([np.stack([src[f, 48 - d[f, c]:48 - d[f, c] + n] for c in range(m)]) for f in range(b)])
    fr = (fr + 0.4 * rng.standard_normal(fr.shape)).astype(np.float32)
    res = pal.gcc_phat_tdoa_batched(torch.from_numpy(fr).cuda(), fs, max_expected_delay=med)
    td = res.tdoa_seconds()[..., 0]
    flags = res.flags.cpu().numpy()
    assert ((flags & 8) != 0).mean() < 0.01
    frd = fr.astype(np.float64)
    pairs = pal.all_pairs(m)
    bad = []
    for f in range(b):
        for p, (i, j) in enumerate(pairs):
            want, _, _ = O.get_time_delays_phat(frd[f, i], frd[f, j], fs, max_expected_delay=med)
            if td[f, p] != want[0]:
                bad.append((f, p, td[f, p], want[0], int(flags[f, p])))
    assert not bad, bad[:5]


def test_float64_ingest_band_limited_signals(pal):
    """Genuine float64 inputs take the float64 ingest (pal_gcc_phat_tdoa_f64): band-passed channels at fs = 44.1 kHz
    (order-5 Butterworth 300-3400 Hz + filtfilt, what main.py:191 hands to the pair loop).  Their stop-band bins sit far
    below the float32 quantisation noise yet PHAT gives them unit weight, so the correlation of the float32-rounded
    channels differs visibly from the reference's; through the float64 path lags are the reference's and the
    correlation agrees to 1e-6 of its maximum -- for the generic length and for 2048-sample frames (n = 4095)."""
    from scipy.signal import butter, filtfilt
    from pyaudiolocalization_b200 import utils as U
    rng = np.random.default_rng(44)
    fs = 44100.0
    b, a = butter(5, [300 / (0.5 * fs), 3400 / (0.5 * fs)], btype="band")
    src = rng.standard_normal(6000)
    x1 = filtfilt(b, a, src[40:40 + 5000] + 0.05 * rng.standard_normal(5000))
    x2 = filtfilt(b, a, src[33:33 + 5000] + 0.05 * rng.standard_normal(5000))
    assert x1.dtype == np.float64
    for med in (None, 0.01):
        want_td, want_corr, _ = O.get_time_delays_phat(x1, x2, fs, max_expected_delay=med)
        td, corr, lags = U.get_time_delays_phat(x1, x2, fs, max_expected_delay=med)
        assert td == want_td
        assert np.abs(corr - want_corr).max() <= 1e-6 * np.abs(want_corr).max()
    c32 = U.phat_correlation(x1.astype(np.float32), x2.astype(np.float32))           # the float32 throughput path
    c64 = U.phat_correlation(x1, x2)
    want = O.phat_correlation(x1, x2)
    e32, e64 = np.abs(c32 - want).max() / np.abs(want).max(), np.abs(c64 - want).max() / np.abs(want).max()
    print(f"band-limited channels: corr error float32 ingest {e32:.2e}, float64 ingest {e64:.2e} (relative to max|corr|)")
    assert e64 <= 1e-6
    # n = 4095: float64 frames [B, M, 2048]
    fr = np.stack([np.stack([filtfilt(b, a, rng.standard_normal(2048)) for _ in range(3)]) for _ in range(2)])
    res = pal.gcc_phat_tdoa_batched(torch.from_numpy(fr).cuda(), 16000.0, max_expected_delay=0.05, return_corr=True)
    assert int(((res.flags & 8) == 0).sum().item()) == 0           # every row went through the float64 kernel
    td = res.tdoa_seconds()[..., 0]
    corr = res.corr.cpu().numpy()
    pairs = pal.all_pairs(3)
    for f in range(2):
        for p, (i, j) in enumerate(pairs):
            w_td, w_corr, _ = O.get_time_delays_phat(fr[f, i], fr[f, j], 16000.0, max_expected_delay=0.05)
            assert td[f, p] == w_td[0]
            assert np.abs(corr[f, p] - w_corr).max() <= 1e-6 * np.abs(w_corr).max()


def test_generic_path_small_workspace_equals_full(pal):
    """Arbitrary-length path with the smallest workspace the library accepts (many rounds; the device-counted float64
    sweep gives way to the host-counted one) against the comfortable workspace: identical lags, peaks and flags."""
    rng = np.random.default_rng(5)
    b, m, n, fs, med = 48, 6, 1500, 16000.0, 0.01
    src = rng.standard_normal((b, n + 64)).astype(np.float32)
    d = rng.integers(0, 48, size=(b, m))
    fr = np.stack([np.stack([src[f, 48 - d[f, c]:48 - d[f, c] + n] for c in range(m)]) for f in range(b)])
    fr = torch.from_numpy((fr + 0.4 * rng.standard_normal(fr.shape)).astype(np.float32)).cuda()
    full, small = pal.gcc_phat.workspace_bytes(b, m, n, m * (m - 1) // 2)
    a = pal.gcc_phat_tdoa_batched(fr, fs, max_expected_delay=med)
    c = pal.gcc_phat_tdoa_batched(fr, fs, max_expected_delay=med, max_workspace_bytes=small)
    assert small < full
    assert torch.equal(a.k_idx, c.k_idx)         # (the flag BITS may differ: the small workspace takes the full-row pick)
    assert torch.allclose(a.peak, c.peak, rtol=0, atol=1e-6) and torch.allclose(a.gmax, c.gmax, rtol=0, atol=1e-6)
    assert int(((a.flags & 8) != 0).sum().item()) > 0          # some rows did go through the float64 sweep


def test_fast_path_odd_microphone_count_and_short_plans(pal):
    """Arbitrary-length fast path with an ODD number of microphones (the last packed forward transform carries a single
    channel; per-channel whitening on: 10 pairs for 5 channels) at three lengths that take the three single-CTA plans
    (n = 1999, 3999, 7999): TDOAs bit-exact against the oracle."""
    rng = np.random.default_rng(123)
    fs, med = 16000.0, 0.02
    for n in (1000, 2000, 4000):
        b, m = 6, 5
        src = rng.standard_normal((b, n + 64)).astype(np.float32)
        d = rng.integers(0, 48, size=(b, m))
        fr = np.stack([np.stack([src[f, 48 - d[f, c]:48 - d[f, c] + n] for c in range(m)]) for f in range(b)])
        fr = (fr + 0.4 * rng.standard_normal(fr.shape)).astype(np.float32)
        res = pal.gcc_phat_tdoa_batched(torch.from_numpy(fr).cuda(), fs, max_expected_delay=med)
        td = res.tdoa_seconds()[..., 0]
        gm = res.gmax.cpu().numpy()
        frd = fr.astype(np.float64)
        for f in range(b):
            for p, (i, j) in enumerate(pal.all_pairs(m)):
                want, corr, _ = O.get_time_delays_phat(frd[f, i], frd[f, j], fs, max_expected_delay=med)
                assert td[f, p] == want[0], (n, f, p)
                assert abs(gm[f, p] - corr.max()) <= CORR_RTOL * corr.max()


@pytest.mark.parametrize("n,m,kw", [(2048, 4, {}), (1500, 5, {}), (4000, 3, {}), (3000, 3, dict(return_corr=True)), (700, 3, {})])
def test_tonal_frames_over_a_deep_noise_floor(pal, n, m, kw):
    """Windowed tones over a -80 dB floor: the stop-band bins sit below the rounding noise of a float32 transform, PHAT
    gives them unit weight all the same, and a float32 correlation is off by 1e-3 of its maximum.  The per-channel
    rounding-noise bound must send such rows to the float64 sweep on every code path: the fused
    n = 4095 kernels, the whitened and the per-pair chirp-z sweeps, full-row picks and the short-transform engine."""
    rng = np.random.default_rng(n + m)
    fs = 16000.0
    t = np.arange(n) / fs
    fr = np.zeros((2, m, n), np.float32)
    for f in range(2):
        for c in range(m):
            d = int(rng.integers(0, 300))
            x = sum(np.sin(2 * np.pi * fq * (t - d / fs)) for fq in (440.0, 1234.5, 0.21 * fs))
            fr[f, c] = x * np.hanning(n) + 1e-4 * rng.standard_normal(n)
    res = pal.gcc_phat_tdoa_batched(torch.from_numpy(fr).cuda(), fs, max_expected_delay=0.05, **kw)
    td = res.tdoa_seconds()[..., 0]
    gm, fl = res.gmax.cpu().numpy(), res.flags.cpu().numpy()
    for f in range(2):
        for p, (i, j) in enumerate(pal.all_pairs(m)):
            want_td, c, _ = O.get_time_delays_phat(fr[f, i].astype(np.float64), fr[f, j].astype(np.float64), fs, max_expected_delay=0.05)
            assert td[f, p] == want_td[0], (f, i, j)
            assert abs(gm[f, p] - c.max()) <= 1e-5 * c.max(), (f, i, j, gm[f, p], c.max())
            assert fl[f, p] & 8          # re-evaluated in float64


def test_pair_table_in_shared_and_in_global_memory(pal):
    """The fused n = 4095 kernel keeps the pair table in shared memory when it fits behind the warp tiles (P <= ~2500)
    and reads it from global memory otherwise.  73 microphones give 2628 pairs (global-memory table); the same rows
    asked for in two halves (shared-memory tables) must come back bit-identical, and a spread of them must match the
    oracle."""
    rng = np.random.default_rng(73)
    fs, med, m = 16000.0, 0.05, 73
    src = rng.standard_normal((2, 2048 + 128)).astype(np.float32)
    d = rng.integers(0, 96, size=(2, m))
    fr = np.stack([np.stack([src[f, 96 - d[f, c]:96 - d[f, c] + 2048] for c in range(m)]) for f in range(2)])
    fr = (fr + 0.3 * rng.standard_normal(fr.shape)).astype(np.float32)
    frd = torch.from_numpy(fr).cuda()
    pairs = np.asarray(pal.all_pairs(m), np.int32)
    assert len(pairs) == 2628
    whole = pal.gcc_phat_tdoa_batched(frd, fs, max_expected_delay=med, pairs=pairs)
    half = len(pairs) // 2
    parts = [pal.gcc_phat_tdoa_batched(frd, fs, max_expected_delay=med, pairs=pairs[s]) for s in (slice(0, half), slice(half, None))]
    for name in ("k_idx", "peak", "gmax"):
        a = getattr(whole, name).cpu().numpy()
        b = np.concatenate([getattr(q, name).cpu().numpy() for q in parts], axis=1)
        assert np.array_equal(a, b), name
    td = whole.tdoa_seconds()[..., 0]
    f64 = fr.astype(np.float64)
    for f in range(2):
        for p in range(0, len(pairs), 41):
            i, j = pairs[p]
            want, corr, _ = O.get_time_delays_phat(f64[f, i], f64[f, j], fs, max_expected_delay=med)
            assert td[f, p] == want[0], (f, p)
            assert abs(whole.gmax[f, p].item() - corr.max()) <= CORR_RTOL * corr.max()
