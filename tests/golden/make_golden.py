#!/usr/bin/env python
"""Generate golden vectors by running the UNMODIFIED reference (read-only at
/root/reference) in the build container.  Only runs where the reference is mounted;
its outputs are committed as small .npz fixtures next to this script.

    python tests/golden/make_golden.py

The reference imports soundfile / resampy / matplotlib at module scope; none of
them is touched on the hot path, so empty stub modules are
installed before importing it.
"""
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("PAL_REFERENCE", "/root/reference")


def import_reference():
    for name in ("soundfile", "resampy", "matplotlib", "matplotlib.pyplot",
                 "mpl_toolkits", "mpl_toolkits.mplot3d"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["mpl_toolkits.mplot3d"].Axes3D = object
    sys.path.insert(0, REF)
    import utils as r_utils                       # noqa: E402
    import signal_processing as r_sp              # noqa: E402
    import main as r_main                         # noqa: E402
    sys.path.pop(0)
    return r_utils, r_sp, r_main


CUSTOM_MATERIALS = {
    "air": {"absorption": 0.01, "freq": 1e-6},
    "wood": {"absorption": 0.05, "freq": 1e-5},
    "metal": {"absorption": 0.1, "freq": 2e-5},
    "glass": {"absorption": 0.07, "freq": 1.5e-5},
}


def shoebox(lx, ly, lz):
    mats = ["wood", "metal", "glass", "wood", "wood", "metal"]
    pl = [[1, 0, 0, 0], [1, 0, 0, -lx], [0, 1, 0, 0], [0, 1, 0, -ly], [0, 0, 1, 0], [0, 0, 1, -lz]]
    return [{"plane": p, "material": m} for p, m in zip(pl, mats)]


def cfg3_frames(seed, frames, mics, n=2048, fs=16000.0):
    """Same generator as bench.py / tests (numpy flavour): noise + speech-like formants,
    integer channel delays, per-channel noise.  float32 like the GPU inputs."""
    rng = np.random.default_rng(seed)
    t = np.arange(n + 64) / fs
    win = np.hanning(n + 64)
    out = np.empty((frames, mics, n), np.float32)
    for f in range(frames):
        speech = (np.sin(2 * np.pi * 800 * t) + 0.8 * np.sin(2 * np.pi * 1150 * t + np.pi / 4)
                  + 0.5 * np.sin(2 * np.pi * 2900 * t + np.pi / 2)) * win
        src = 0.5 * rng.standard_normal(n + 64) + 0.5 * speech
        d = rng.integers(0, 40, size=mics)
        for m in range(mics):
            out[f, m] = (src[40 - d[m]: 40 - d[m] + n] + 0.3 * rng.standard_normal(n)).astype(np.float32)
    return out


def main():
    ru, rsp, rmain = import_reference()
    g = {}

    # --- scalars / geometry -------------------------------------------------
    g["sos_20_50"] = ru.speed_of_sound(20, 50)
    g["sos_clamped"] = ru.speed_of_sound(80, 120)
    g["reflect"] = ru.reflect_point_across_plane([0.5, 0.25, 1.5], [1, 2, -1, -3])
    from materials import material_properties as stock
    g["att_direct_cfg1"] = ru.calculate_attenuation(ru.distance([0.5, 0.5, 0.5], [0, 0, 0]), "air", 1000, stock)

    # --- image sources --------------------------------------------------------
    rng = np.random.default_rng(0)
    mics8 = rng.uniform([1, 1, 0.5], [5, 4, 2.5], size=(8, 3))
    src = np.array([2.2, 3.1, 1.4])
    counts = []
    for order in range(1, 7):
        im = ru.generate_image_sources_iterative(src, shoebox(6, 5, 3), order, 1000.0, CUSTOM_MATERIALS,
                                                 mics8, absorption_threshold=-1.0)
        counts.append(len(im))
    g["img_counts_thr_neg1"] = np.array(counts)
    im = ru.generate_image_sources_iterative(src, shoebox(6, 5, 3), 3, 1000.0, CUSTOM_MATERIALS, mics8,
                                             absorption_threshold=0.01)
    g["img_mics"] = mics8
    g["img_src"] = src
    g["img_pos_o3"] = np.array([i["source"] for i in im])
    names = sorted(CUSTOM_MATERIALS)
    g["img_mat_o3"] = np.array([names.index(i["material"]) for i in im])
    g["img_cfg1_count"] = len(ru.generate_image_sources_iterative(
        [0.5, 0.5, 0.5], rmain.config["reflective_planes"], 3, 1000, stock,
        np.array(rmain.config["mic_positions"]), 0.01))

    # --- fractional delay / render ---------------------------------------------
    fs = 16000.0
    base = rsp.generate_signal("chirp", fs, 0.25, 500)
    g["chirp_16k_025_500"] = base
    g["frac_delay"] = rsp.fractional_delay(np.pad(base, (0, 300)), 0.00731, fs)
    mics4 = mics8[:4]
    sig = rmain.simulate_signals_with_multipath(src, mics4, fs, 343.62, duration=0.25, signal_type="chirp",
                                                freq=500, reflective_planes=shoebox(6, 5, 3),
                                                material_properties=CUSTOM_MATERIALS, max_reflections=2,
                                                absorption_threshold=0.01)
    g["render_o2_4mics"] = np.array(sig)
    x = np.linspace(-2, 2, 257) ** 3
    g["compress_in"] = x
    g["compress_out"] = rsp.dynamic_range_compression(x.copy())

    # --- GCC-PHAT golden (4) of SURVEY §8c ---------------------------------------
    r0 = np.random.default_rng(0)
    xx = r0.standard_normal(2148)
    s1, s2 = xx[100:2148], xx[93:2141]
    corr = ru.phat_correlation(s1, s2)
    g["g4_argmax"] = int(np.argmax(corr))
    g["g4_corr"] = corr
    for tag, med in (("none", None), ("0p05", 0.05), ("0p01", 0.01)):
        td, _, _ = ru.get_time_delays_phat(s1, s2, 16000, max_expected_delay=med)
        g[f"g4_td_{tag}"] = np.array(td)
    td, corr0, _ = ru.get_time_delays_phat(np.zeros(256), np.zeros(256), 16000, max_expected_delay=0.05)
    g["zero_td"] = np.array(td)
    g["zero_max"] = float(np.max(corr0))

    # --- cfg3-shaped frames: every pair of 8 mics, two frames --------------------
    fr = cfg3_frames(1234, 2, 8)
    g["cfg3_frames"] = fr
    tds, kmax = [], []
    for f in range(fr.shape[0]):
        for i in range(8):
            for j in range(i + 1, 8):
                td, corr, _ = ru.get_time_delays_phat(fr[f, i].astype(np.float64), fr[f, j].astype(np.float64),
                                                      16000.0, num_peaks=1, max_expected_delay=0.05)
                tds.append(td[0])
                kmax.append(np.max(corr))
    g["cfg3_td"] = np.array(tds)
    g["cfg3_gmax"] = np.array(kmax)

    # --- fuzz of get_time_delays_phat control flow (short rows, many option combos) ----
    rf = np.random.default_rng(7)
    rows = []
    sigs = []
    for case in range(160):
        n1 = int(rf.integers(40, 400))
        n2 = int(rf.integers(40, 400)) if case % 3 else n1
        fs_c = float(rf.choice([8000, 16000, 44100, 48000]))
        method = ["median", "adaptive", "other"][case % 3]
        mult = float(rf.choice([0.5, 1.0, 3.0, 8.0]))
        med = [None, 0.05, 0.01, 0.003, 0.0005][case % 5]
        npk = int(rf.choice([1, 1, 2, 3]))
        a = rf.standard_normal(n1)
        b = np.roll(a, int(rf.integers(0, 9)))[:n2] if n2 <= n1 else rf.standard_normal(n2)
        b = b + 0.2 * rf.standard_normal(len(b))
        if case % 11 == 0:
            a = np.round(a * 2) / 2
            b = np.round(b * 2) / 2
        if case % 37 == 0:
            a = np.zeros(n1)
        td, _, _ = ru.get_time_delays_phat(a, b, fs_c, num_peaks=npk, threshold_method=method,
                                           threshold_multiplier=mult, max_expected_delay=med)
        rows.append((n1, n2, fs_c, ["median", "adaptive", "other"].index(method), mult,
                     -1.0 if med is None else med, npk, len(td)))
        sigs.append((a, b, np.array(td)))
    g["fuzz_meta"] = np.array(rows, dtype=np.float64)
    for i, (a, b, td) in enumerate(sigs):
        g[f"fuzz_a{i}"] = a
        g[f"fuzz_b{i}"] = b
        g[f"fuzz_td{i}"] = td

    np.savez_compressed(os.path.join(HERE, "reference_vectors.npz"), **g)
    print("wrote", os.path.join(HERE, "reference_vectors.npz"), len(g), "arrays")


if __name__ == "__main__":
    main()
