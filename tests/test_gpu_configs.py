"""GPU parity on the workloads BASELINE.json names, as SURVEY.md section 8d states them (run with -m gpu):
cfg2 (1 s @ 44.1 kHz chirp + noise, n = 88 199), cfg3 (64-frame subset, every pair), cfg4 (64 mics, order-6 image
sources; n = 95 999 on rendered channels), cfg5 (random rooms: render -> GCC-PHAT chain).  Everything goes through the
C ABI.  Bars: lag indices / TDOAs bit-exact, correlation maxima within 1e-4 relative, rendered channels within 1e-5."""
import numpy as np
import pytest

from oracle import pal_oracle as O
from tests import _pool
from tests.golden.make_golden import CUSTOM_MATERIALS, shoebox

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

CORR_RTOL = 1e-4
RENDER_ATOL = 1e-5


@pytest.fixture(scope="module")
def pal():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import pyaudiolocalization_b200 as p
    return p


def test_cfg2_chirp_scenes_n88199(pal):
    """cfg2 generator exactly as SURVEY 8d: base = generate_signal('chirp', 44100, 1.0, 1000) (1 -> 5 kHz), channel m =
    base delayed by d ~ U{0..63} samples (leading zeros, tail cut) + 0.05 N(0,1), float32; 4 mics, 6 pairs,
    n = 88 199, max_expected_delay = 0.05 s (W = 2205, D = 44)."""
    fs, n, m, b = 44100.0, 44100, 4, 6
    base = O.generate_signal("chirp", fs, 1.0, 1000)
    rng = np.random.default_rng(2000)
    fr = np.zeros((b, m, n), np.float32)
    for s in range(b):
        for c in range(m):
            d = int(rng.integers(0, 64))
            fr[s, c, d:] = base[:n - d]
            fr[s, c] += (0.05 * rng.standard_normal(n)).astype(np.float32)
    res = pal.gcc_phat_tdoa_batched(torch.from_numpy(fr).cuda(), fs, max_expected_delay=0.05)
    td = res.tdoa_seconds()[..., 0]
    gm = res.gmax.cpu().numpy()
    want_td, want_gm = _pool.oracle_pairs(fr, fs, 0.05)
    assert np.array_equal(td, want_td), np.argwhere(td != want_td)[:5]
    assert np.max(np.abs(gm - want_gm) / np.abs(want_gm)) <= CORR_RTOL
    assert pal.window_half_width(n, n, fs, 0.05) == 2205 and pal.peak_distance(fs) == 44


def test_cfg3_parity_subset_64_frames_every_pair(pal):
    """SURVEY 8d parity subset at the benchmark shape and size: all 496 pairs of the first 64 frames of the cfg3
    workload (31 744 rows; generator and seed of bench.py): TDOA bit-exact, max(corr) within 1e-4."""
    from pyaudiolocalization_b200 import synth
    fr = synth.cfg3_frames(64, mics=32, seed=3000)
    res = pal.gcc_phat_tdoa_batched(fr, 16000.0, max_expected_delay=0.05)
    td = res.tdoa_seconds()[..., 0]
    gm = res.gmax.cpu().numpy()
    want_td, want_gm = _pool.oracle_pairs(fr.cpu().numpy(), 16000.0, 0.05)
    assert int((td != want_td).sum()) == 0
    assert np.max(np.abs(gm - want_gm) / np.abs(want_gm)) <= CORR_RTOL
    refined = int(((res.flags & 8) != 0).sum().item())
    print("cfg3 subset: 31744 rows,", refined, "through the float64 kernel")
    assert refined > 0          # the subset is large enough to exercise the float64 path


def test_cfg4_image_sources_64_mics_order_6(pal):
    """cfg4 geometry: shoebox 6 x 5 x 3 m with 6 material planes, 64 mics U([1,1,.5],[5,4,2.5]) seed 0, sources
    U([.5,.5,.3],[5.5,4.5,2.7]) seed 1, max_reflections = 6: bit-exact positions / order / materials; 376 images when
    nothing is pruned."""
    from pyaudiolocalization_b200 import scene
    mics = np.random.default_rng(0).uniform([1, 1, 0.5], [5, 4, 2.5], size=(64, 3))
    srcs = np.random.default_rng(1).uniform([0.5, 0.5, 0.3], [5.5, 4.5, 2.7], size=(1024, 3))[:6]
    planes = shoebox(6, 5, 3)
    for thr, expect in ((0.01, None), (-1.0, 376)):
        pos, mat, cnt, table = scene.image_sources_batched(srcs, planes, 6, 1000.0, CUSTOM_MATERIALS, mics, thr)
        pos, mat, cnt = pos.cpu().numpy(), mat.cpu().numpy(), cnt.cpu().numpy()
        for s in range(len(srcs)):
            im = O.generate_image_sources_iterative(srcs[s], planes, 6, 1000.0, CUSTOM_MATERIALS, mics, thr)
            assert cnt[s] == len(im) and (expect is None or cnt[s] == expect)
            assert np.array_equal(pos[s, :cnt[s]], np.array([i["source"] for i in im]).reshape(-1, 3))
            assert [table.names[k] for k in mat[s, :cnt[s]]] == [i["material"] for i in im]


def test_cfg4_rendered_pairs_n95999(pal):
    """cfg4 stage 1 -> stage 2 on one source: 1 s @ 48 kHz chirp rendered with order-6 reflections into 4 of the 64
    microphones (rendered channels vs the oracle's transfer-function restatement <= 1e-5), then every pair of the
    GPU-rendered float32 channels through GCC-PHAT at n = 95 999: TDOA bit-exact vs the oracle on those channels."""
    from pyaudiolocalization_b200 import main as M
    fs, dur, freq, c = 48000.0, 1.0, 1000.0, 343.62
    mics = np.random.default_rng(0).uniform([1, 1, 0.5], [5, 4, 2.5], size=(64, 3))[[0, 17, 40, 63]]
    src = np.random.default_rng(1).uniform([0.5, 0.5, 0.3], [5.5, 4.5, 2.7], size=(1024, 3))[0]
    planes = shoebox(6, 5, 3)
    sig = M.simulate_scenes_batched(src[None], mics, fs, c, dur, "chirp", freq, planes, CUSTOM_MATERIALS, 6, 0.01)
    assert tuple(sig.shape) == (1, 4, 48000)
    imgs = O.generate_image_sources_iterative(src, planes, 6, freq, CUSTOM_MATERIALS, mics, 0.01)
    assert len(imgs) > 200
    tau, gain, total = O.path_table_restated(src, imgs, mics, fs, c, dur, freq, CUSTOM_MATERIALS)
    want = O.render_rows_restated(O.generate_signal("chirp", fs, dur, freq), tau, gain, total, fs, int(dur * fs))
    got = sig[0].cpu().numpy()
    err = np.abs(got - want).max()
    print(f"cfg4 render ({len(imgs)} images, N = {total}): max abs error {err:.2e}")
    assert err <= RENDER_ATOL
    res = pal.gcc_phat_tdoa_batched(sig, fs, max_expected_delay=0.05)
    want_td, want_gm = _pool.oracle_pairs(got[None], fs, 0.05)
    assert np.array_equal(res.tdoa_seconds()[..., 0], want_td)
    assert np.max(np.abs(res.gmax.cpu().numpy() - want_gm) / np.abs(want_gm)) <= CORR_RTOL


def test_cfg5_chain_render_then_gcc(pal):
    """cfg5 chain on 16 random rooms x 8 mics (0.25 s @ 16 kHz chirp, order 3): rendered channels vs the oracle
    (port of main.py:66-124 on 3 scenes, transfer-function restatement on all 16) <= 1e-5, then the TDOAs of the
    GPU-rendered float32 channels bit-exact vs the oracle's GCC-PHAT on those same channels (28 pairs, n = 7999)."""
    from pyaudiolocalization_b200 import sweep
    cfg = sweep.SweepConfig()
    n_sc = 16
    src, mic, pl = sweep.random_shoebox_scenes(n_sc, cfg.mics, 5000)
    sw = sweep.SceneSweep(cfg, n_sc, chunk=6, keep_signals=n_sc)       # three chunks: the software-pipelined path
    k = sw.step(src, mic, pl)
    torch.cuda.synchronize()
    sig = sw.signals.cpu().numpy()
    base = O.generate_signal(cfg.signal_type, cfg.fs, cfg.duration, cfg.freq)
    worst = 0.0
    for s in range(n_sc):
        planes = sweep.planes_as_dicts(pl[s])
        imgs = O.generate_image_sources_iterative(src[s], planes, cfg.max_reflections, cfg.freq, sweep.SWEEP_MATERIALS, mic[s],
                                                  cfg.absorption_threshold)
        tau, gain, total = O.path_table_restated(src[s], imgs, mic[s], cfg.fs, cfg.c, cfg.duration, cfg.freq, sweep.SWEEP_MATERIALS)
        want = O.render_rows_restated(base, tau, gain, total, cfg.fs, cfg.samples)
        if s < 3:
            port = np.array(O.simulate_signals_with_multipath(src[s], mic[s], cfg.fs, cfg.c, duration=cfg.duration,
                                                              signal_type=cfg.signal_type, freq=cfg.freq, reflective_planes=planes,
                                                              material_properties=sweep.SWEEP_MATERIALS,
                                                              max_reflections=cfg.max_reflections,
                                                              absorption_threshold=cfg.absorption_threshold))
            assert np.abs(port - want).max() <= 1e-9         # the restatement IS the reference's renderer
        worst = max(worst, float(np.abs(sig[s] - want).max()))
    print(f"cfg5 chain: render max abs error over {n_sc} scenes {worst:.2e}")
    assert worst <= RENDER_ATOL
    td = pal.shard.tdoa_seconds_from_indices(k, cfg.samples, float(cfg.fs))[..., 0]
    want_td, _ = _pool.oracle_pairs(sig, float(cfg.fs), cfg.max_expected_delay)
    assert np.array_equal(td, want_td), np.argwhere(td != want_td)[:5]
