"""Secondary measurements on the other BASELINE.json configurations (the bench.py line is cfg3).
Each prints one JSON line; sizes are bounded sub-samples of the named configuration.

    python tools/bench_configs.py [cfg2 cfg4 cfg5 img] [--small]
"""
import json
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
import pyaudiolocalization_b200 as pal
from pyaudiolocalization_b200 import main as pmain, scene
from pyaudiolocalization_b200.signal_processing import generate_signal

MATS = {"air": {"absorption": 0.01, "freq": 1e-6}, "wood": {"absorption": 0.05, "freq": 1e-5},
        "metal": {"absorption": 0.1, "freq": 2e-5}, "glass": {"absorption": 0.07, "freq": 1.5e-5}}


def shoebox(lx, ly, lz):
    m = ["wood", "metal", "glass", "wood", "wood", "metal"]
    pl = [[1, 0, 0, 0], [1, 0, 0, -lx], [0, 1, 0, 0], [0, 1, 0, -ly], [0, 0, 1, 0], [0, 0, 1, -lz]]
    return [{"plane": p, "material": mm} for p, mm in zip(pl, m)]


def timed(fn, reps=3, warm=1):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t) / reps


def stage2(name, B, M, N, fs, med, seed):
    g = torch.Generator(device="cuda"); g.manual_seed(seed)
    base = torch.as_tensor(generate_signal("chirp", fs, N / fs, 1000.0 if fs > 20000 else 500.0)[:N], dtype=torch.float32, device="cuda")
    d = torch.randint(0, 64, (B, M), generator=g, device="cuda")
    idx = torch.arange(N, device="cuda")[None, None, :] - d[:, :, None]
    fr = torch.where(idx >= 0, base[idx.clamp(min=0)], torch.zeros((), device="cuda"))
    fr = fr + 0.05 * torch.randn((B, M, N), generator=g, device="cuda")
    P = M * (M - 1) // 2
    full, small = pal.gcc_phat.workspace_bytes(B, M, N, P)
    ws = torch.empty(min(full, 8 << 30) + 256, dtype=torch.uint8, device="cuda")
    sec = timed(lambda: pal.gcc_phat_tdoa_batched(fr, fs, med, workspace=ws), reps=2)
    r = pal.gcc_phat_tdoa_batched(fr, fs, med, workspace=ws)
    refined = float(((r.flags & 8) != 0).float().mean().item())
    alg = 20 * M * N + 16 * P * N + 16 * P
    print(json.dumps({"config": name, "stage": "gcc_phat_tdoa", "units": B, "mics": M, "samples": N, "n_fft": 2 * N - 1,
                      "ms": sec * 1e3, "scenes_per_s": B / sec, "pair_corr_per_s": B * P / sec,
                      "algorithmic_GBps": alg * B / sec / 1e9, "refined_row_fraction": refined}), flush=True)


def cfg4_full():
    """BASELINE cfg4 at its stated size: all 1024 source positions rendered (64 mics, order 6, 1 s @ 48 kHz; chunks of 64
    sources), then stage 2 on 8 of the rendered sources (8 x 2016 pairs, n = 95 999)."""
    rng0, rng1 = np.random.default_rng(0), np.random.default_rng(1)
    mics = rng0.uniform([1, 1, 0.5], [5, 4, 2.5], size=(64, 3))
    srcs = rng1.uniform([0.5, 0.5, 0.3], [5.5, 4.5, 2.7], size=(1024, 3))
    planes = shoebox(6, 5, 3)
    base = torch.as_tensor(generate_signal("chirp", 48000, 1.0, 1000).astype(np.float32)).cuda()
    cache = scene.RenderPlanCache()

    def render_all(keep=0):
        kept = None
        for s0 in range(0, 1024, 64):
            out = pmain.simulate_scenes_batched(srcs[s0:s0 + 64], mics, 48000, 343.62, 1.0, "chirp", 1000, planes, MATS, 6, 0.01,
                                                base_signal=base, plan_cache=cache)
            if s0 == 0 and keep:
                kept = out[:keep].clone()
        return kept
    render_all()
    sec = timed(render_all, reps=1, warm=0)
    print(json.dumps({"config": "cfg4 (full size)", "stage": "render", "mics": 64, "order": 6, "fs": 48000, "sources": 1024,
                      "seconds": sec, "sources_per_s": 1024 / sec, "rendered_samples_per_s": 1024 * 64 * 48000 / sec,
                      "reference_estimate": "about 2000 s per source position on one core (SURVEY section 6)"}), flush=True)
    fr = render_all(keep=8)
    P = 64 * 63 // 2
    sec = timed(lambda: pal.gcc_phat_tdoa_batched(fr, 48000.0, 0.05), reps=2)
    r = pal.gcc_phat_tdoa_batched(fr, 48000.0, 0.05)
    print(json.dumps({"config": "cfg4 (full size)", "stage": "gcc_phat_tdoa on 8 rendered sources", "units": 8, "mics": 64, "pairs": P,
                      "samples": 48000, "n_fft": 95999, "ms": sec * 1e3, "pair_corr_per_s": 8 * P / sec,
                      "refined_row_fraction": float(((r.flags & 8) != 0).float().mean().item())}), flush=True)


def cfg4(small):
    rng0, rng1 = np.random.default_rng(0), np.random.default_rng(1)
    mics = rng0.uniform([1, 1, 0.5], [5, 4, 2.5], size=(64, 3))
    srcs = rng1.uniform([0.5, 0.5, 0.3], [5.5, 4.5, 2.7], size=(1024, 3))
    planes = shoebox(6, 5, 3)
    n_src = 8 if small else 64
    sec1 = timed(lambda: pmain.simulate_signals_device(srcs[0], mics, 48000, 343.62, 1.0, "chirp", 1000, planes, MATS, 6, 0.01), reps=2)
    sec = timed(lambda: pmain.simulate_scenes_batched(srcs[:n_src], mics, 48000, 343.62, 1.0, "chirp", 1000, planes, MATS, 6, 0.01), reps=2)
    print(json.dumps({"config": "cfg4", "stage": "render (image sources + path table + transfer + inverse DFT + normalise/compress)",
                      "mics": 64, "order": 6, "fs": 48000, "sources": n_src, "batched_ms": sec * 1e3,
                      "sources_per_s_batched": n_src / sec, "sources_per_s_per_scene_api": 1 / sec1,
                      "rendered_samples_per_s": n_src * 64 * 48000 / sec}), flush=True)
    fr = pmain.simulate_scenes_batched(srcs[:2], mics, 48000, 343.62, 1.0, "chirp", 1000, planes, MATS, 6, 0.01)
    P = 64 * 63 // 2
    sec = timed(lambda: pal.gcc_phat_tdoa_batched(fr, 48000.0, 0.05), reps=1)
    print(json.dumps({"config": "cfg4", "stage": "gcc_phat_tdoa", "units": 2, "mics": 64, "samples": 48000, "n_fft": 95999,
                      "ms": sec * 1e3, "pair_corr_per_s": 2 * P / sec}), flush=True)


def cfg5_render(small):
    rng = np.random.default_rng(5000)
    n_sc = 256 if small else 4096
    rooms, micl, srcl = [], [], []
    for _ in range(n_sc):
        dims = rng.uniform([3, 3, 2.5], [10, 8, 4])
        rooms.append(shoebox(*dims))
        micl.append(rng.uniform([0.3, 0.3, 0.3], dims - 0.3, size=(8, 3)))
        srcl.append(rng.uniform([0.3, 0.3, 0.3], dims - 0.3))
    micl, srcl = np.array(micl), np.array(srcl)

    def per_scene():
        for s in range(16):
            pmain.simulate_signals_device(srcl[s], micl[s], 16000, 343.62, 0.25, "chirp", 500, rooms[s], MATS, 3, 0.01)
    sec1 = timed(per_scene, reps=2) / 16
    sec = timed(lambda: pmain.simulate_scenes_batched(srcl, micl, 16000, 343.62, 0.25, "chirp", 500, rooms, MATS, 3, 0.01), reps=2)
    # the same with the renderer's per-length tables kept between calls (steady state of a sweep)
    base_dev = torch.as_tensor(generate_signal("chirp", 16000, 0.25, 500).astype(np.float32)).cuda()
    cache = scene.RenderPlanCache()
    secc = timed(lambda: pmain.simulate_scenes_batched(srcl, micl, 16000, 343.62, 0.25, "chirp", 500, rooms, MATS, 3, 0.01,
                                                       base_signal=base_dev, plan_cache=cache), reps=2)
    print(json.dumps({"config": "cfg5", "stage": "render", "mics": 8, "order": 3, "fs": 16000, "scenes": n_sc,
                      "batched_ms": sec * 1e3, "scenes_per_s_batched": n_sc / sec, "scenes_per_s_per_scene_api": 1 / sec1,
                      "batched_ms_plan_cache": secc * 1e3, "scenes_per_s_batched_plan_cache": n_sc / secc}), flush=True)
    sig = pmain.simulate_scenes_batched(srcl, micl, 16000, 343.62, 0.25, "chirp", 500, rooms, MATS, 3, 0.01)
    P = 28
    sec2 = timed(lambda: pal.gcc_phat_tdoa_batched(sig, 16000.0, 0.05), reps=2)
    print(json.dumps({"config": "cfg5", "stage": "render -> gcc_phat_tdoa (rendered scenes fed to stage 2)", "scenes": n_sc,
                      "gcc_ms": sec2 * 1e3, "scenes_per_s_gcc": n_sc / sec2, "scenes_per_s_both": n_sc / (sec + sec2)}), flush=True)


def img(small):
    rng = np.random.default_rng(1)
    mics = np.random.default_rng(0).uniform([1, 1, 0.5], [5, 4, 2.5], size=(64, 3))
    srcs = rng.uniform([0.5, 0.5, 0.3], [5.5, 4.5, 2.7], size=(1024, 3))
    sec = timed(lambda: scene.image_sources_batched(srcs, shoebox(6, 5, 3), 6, 1000.0, MATS, mics, 0.01), reps=2)
    pos, mat, cnt, _ = scene.image_sources_batched(srcs, shoebox(6, 5, 3), 6, 1000.0, MATS, mics, 0.01)
    print(json.dumps({"config": "cfg4", "stage": "image_sources", "scenes": 1024, "order": 6, "mics": 64, "ms": sec * 1e3,
                      "scenes_per_s": 1024 / sec, "mean_images": float(cnt.float().mean().item())}), flush=True)
    n = 16384 if small else 65536
    mics8 = np.random.default_rng(2).uniform([0.3, 0.3, 0.3], [2.7, 2.7, 2.2], size=(8, 3))
    srcs = np.random.default_rng(3).uniform([0.3, 0.3, 0.3], [2.7, 2.7, 2.2], size=(n, 3))
    sec = timed(lambda: scene.image_sources_batched(srcs, shoebox(3, 3, 2.5), 3, 500.0, MATS, mics8, 0.01), reps=2)
    print(json.dumps({"config": "cfg5", "stage": "image_sources", "scenes": n, "order": 3, "mics": 8, "ms": sec * 1e3,
                      "scenes_per_s": n / sec}), flush=True)


def filt(small):
    from pyaudiolocalization_b200 import filters
    frames = 256 if small else 2048
    fr = pal.synth.cfg3_frames(frames, 32, seed=3000)
    b, a = filters.design_butter_bandpass(16000.0)
    sec = timed(lambda: filters.filtfilt_batched(fr, b, a), reps=2)
    rows = frames * 32
    print(json.dumps({"config": "cfg3", "stage": "channel filter (noise_reduction butterworth, filtfilt order 5 band-pass, float64 arithmetic)",
                      "rows": rows, "samples": 2048, "ms": sec * 1e3, "channels_per_s": rows / sec,
                      "samples_per_s": rows * 2048 / sec, "GBps_io": rows * 2048 * 8 / sec / 1e9}), flush=True)


if __name__ == "__main__":
    args = [a for a in sys.argv[1:] if not a.startswith("--")] or ["img", "cfg5", "cfg2", "cfg4", "filter"]
    small = "--small" in sys.argv
    full = "--full" in sys.argv          # the sizes BASELINE.json states: cfg2 4096 scenes, cfg4 1024 sources
    for a in args:
        if a == "cfg2":
            stage2("cfg2" + (" (full size)" if full else ""), 4096 if full else (64 if small else 256), 4, 44100, 44100.0, 0.05, 2000)
        elif a == "cfg4" and full:
            cfg4_full()
        elif a == "cfg5":
            stage2("cfg5", 1024 if small else 8192, 8, 4000, 16000.0, 0.05, 5000)
            cfg5_render(small)
        elif a == "cfg4":
            cfg4(small)
        elif a == "img":
            img(small)
        elif a == "filter":
            filt(small)
