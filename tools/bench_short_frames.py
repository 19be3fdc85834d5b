"""Arbitrary-length GCC-PHAT path on short frames (not 2048 samples): scenes/s for a few frame lengths.
PAL_FUSED=0 disables the single-block short-transform path (A/B)."""
import json
import sys
import time

import torch

sys.path.insert(0, ".")
import pyaudiolocalization_b200 as pal

for ns, b in ((500, 16384), (1000, 16384), (2000, 8192), (4000, 8192)):
    g = torch.Generator(device="cuda"); g.manual_seed(ns)
    fr = torch.randn((b, 8, ns), generator=g, device="cuda")
    fr[:, 1:] = 0.5 * fr[:, :1] + 0.5 * fr[:, 1:]
    for _ in range(2):
        r = pal.gcc_phat_tdoa_batched(fr, 16000.0, 0.01)
    torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(3):
        r = pal.gcc_phat_tdoa_batched(fr, 16000.0, 0.01)
    torch.cuda.synchronize()
    sec = (time.perf_counter() - t) / 3
    print(json.dumps({"samples": ns, "n_fft": 2 * ns - 1, "frames": b, "ms": sec * 1e3, "frames_per_s": b / sec,
                      "pair_corr_per_s": b * 28 / sec}), flush=True)
