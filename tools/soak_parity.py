"""Randomised parity soak of the GCC-PHAT / TDOA path against the oracle (test infrastructure; run on a GPU box):

    python tools/soak_parity.py [--cases 120] [--seed 1] [--budget-s 240]

Every case draws a frame length, a microphone count, a sample rate, a window (max_expected_delay), a threshold method
and a signal family (noise, delayed copies of one source with noise, chirps, sparse clicks, a dead or very quiet
channel, band-limited tones, int16-quantised audio), runs `gcc_phat_tdoa_batched` and compares EVERY pair of every
frame with `oracle.pal_oracle.get_time_delays_phat` on the float64 view of the same float32 samples: TDOAs must be
identical, max(corr) within 1e-4.  One JSON line summarises the run; mismatches are listed (and the exit code is 1).
The oracle side runs in a process pool forked before this process touches CUDA.
"""
import argparse
import json
import multiprocessing as mp
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

LENGTHS = [257, 600, 1024, 1500, 2048, 2048, 2048, 3000, 4000, 4096, 7000, 8192, 12000, 16000, 22050, 44100, 48000]
RATES = [8000.0, 16000.0, 16000.0, 22050.0, 44100.0, 48000.0]
DELAYS = [None, 0.0004, 0.002, 0.01, 0.05, 0.05, 0.3]
FAMILIES = ["noise", "delayed", "delayed", "chirp", "clicks", "dead", "quiet", "tones", "pcm16", "strong_echo"]


def make_case(idx, seed):
    rng = np.random.default_rng([seed, idx])
    n = int(rng.choice(LENGTHS))
    m = int(rng.integers(2, 7))
    b = int(rng.integers(1, 4)) if n <= 16000 else 1
    fs = float(rng.choice(RATES))
    med = DELAYS[int(rng.integers(len(DELAYS)))]
    fam = FAMILIES[int(rng.integers(len(FAMILIES)))]
    kw = {}
    if rng.random() < 0.25:
        kw["threshold_method"] = "adaptive"
        kw["threshold_multiplier"] = float(rng.choice([2.0, 3.0, 5.0]))
    elif rng.random() < 0.2:
        kw["threshold_multiplier"] = float(rng.choice([1.0, 8.0]))
    if rng.random() < 0.15:
        kw["num_peaks"] = int(rng.integers(2, 4))
    fr = np.zeros((b, m, n), np.float64)
    t = np.arange(n) / fs
    for f in range(b):
        src = rng.standard_normal(n + 400)
        for c in range(m):
            d = int(rng.integers(0, 300))
            if fam == "noise":
                x = rng.standard_normal(n)
            elif fam == "delayed":
                x = src[d:d + n] + float(rng.choice([0.01, 0.3, 1.0])) * rng.standard_normal(n)
            elif fam == "chirp":
                f0, f1 = 200.0 + 100 * c, 0.4 * fs
                x = np.sin(2 * np.pi * (f0 * t + 0.5 * (f1 - f0) / t[-1] * t * t)) + 0.05 * rng.standard_normal(n)
                x = np.roll(x, d)
            elif fam == "clicks":
                x = np.zeros(n)
                pos = rng.integers(0, n, size=5)
                x[pos] = rng.standard_normal(5)
                x = np.roll(x, d) + 1e-3 * rng.standard_normal(n)
            elif fam == "dead":
                x = src[d:d + n] if c != 1 else np.zeros(n)
            elif fam == "quiet":
                x = (1e-5 if c % 2 else 1.0) * (src[d:d + n] + 0.1 * rng.standard_normal(n))
            elif fam == "tones":
                x = sum(np.sin(2 * np.pi * fq * (t - d / fs)) for fq in (440.0, 1234.5, 0.21 * fs))
                x = x * np.hanning(n) + 1e-4 * rng.standard_normal(n)
            elif fam == "pcm16":
                x = np.round(3000 * (src[d:d + n] + 0.2 * rng.standard_normal(n))) / 32768.0
            else:  # strong_echo: two copies of the source of equal strength -> near ties
                x = src[d:d + n] + src[d + 37:d + 37 + n] + 0.01 * rng.standard_normal(n)
            fr[f, c] = x
    return {"idx": idx, "family": fam, "n": n, "m": m, "b": b, "fs": fs, "med": med, "kw": kw,
            "frames": fr.astype(np.float32)}


def oracle_case(case):
    from oracle import pal_oracle as O
    fr = case["frames"].astype(np.float64)
    out = []
    for f in range(fr.shape[0]):
        for i in range(fr.shape[1]):
            for j in range(i + 1, fr.shape[1]):
                td, corr, _ = O.get_time_delays_phat(fr[f, i], fr[f, j], case["fs"], max_expected_delay=case["med"], **case["kw"])
                out.append((np.asarray(td, np.float64), float(corr.max())))
    return case["idx"], out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cases", type=int, default=120)
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--budget-s", type=float, default=240.0)
    ap.add_argument("--procs", type=int, default=0)
    a = ap.parse_args()
    cases = [make_case(i, a.seed) for i in range(a.cases)]
    procs = a.procs or min(32, os.cpu_count() or 1)
    pool = mp.get_context("fork").Pool(procs)           # before CUDA is initialised
    try:
        pending = pool.map_async(oracle_case, cases, chunksize=1)
        import torch
        import pyaudiolocalization_b200 as pal
        got = {}
        t0 = time.time()
        for c in cases:
            k = c["kw"].get("num_peaks", 1)
            res = pal.gcc_phat_tdoa_batched(torch.from_numpy(c["frames"]).cuda(), c["fs"], max_expected_delay=c["med"], **c["kw"])
            got[c["idx"]] = (res.tdoa_seconds().reshape(-1, k), res.k_count.cpu().numpy().reshape(-1),
                             res.gmax.cpu().numpy().reshape(-1), res.flags.cpu().numpy().reshape(-1))
        gpu_s = time.time() - t0
        want = dict(pending.get(timeout=a.budget_s))
    finally:
        pool.terminate()
        pool.join()
    rows = bad = refined = 0
    worst_gm = 0.0
    fam_rows = {}
    listing = []
    for c in cases:
        td, cnt, gm, fl = got[c["idx"]]
        for r, (wtd, wgm) in enumerate(want[c["idx"]]):
            rows += 1
            fam_rows[c["family"]] = fam_rows.get(c["family"], 0) + 1
            refined += int((fl[r] & 8) != 0)
            ok = cnt[r] == len(wtd) and np.array_equal(td[r, :cnt[r]], wtd)
            if wgm > 0:
                e = float(abs(float(gm[r]) - wgm) / wgm)
                worst_gm = max(worst_gm, e)
                ok = ok and e <= 1e-4
            else:
                ok = ok and gm[r] == 0
            if not ok:
                bad += 1
                if len(listing) < 20:
                    listing.append({k: c[k] for k in ("idx", "family", "n", "m", "b", "fs", "med", "kw")} |
                                   {"row": r, "got": td[r, :cnt[r]].tolist(), "want": wtd.tolist(), "gmax": [float(gm[r]), wgm],
                                    "flags": int(fl[r])})
    print(json.dumps({"tool": "soak_parity", "seed": a.seed, "cases": len(cases), "rows": rows, "mismatches": bad,
                      "refined_rows": refined, "max_rel_gmax_err": float(worst_gm), "rows_by_family": fam_rows,
                      "gpu_seconds": round(gpu_s, 2)}))
    for item in listing:
        print("MISMATCH", json.dumps(item))
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
