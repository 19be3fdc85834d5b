"""Seeded synthetic workloads of BASELINE.json, generated directly in device memory
(cfg3 = 32 mics x 2048-sample frames of noise + speech-like signal)."""
from __future__ import annotations

import math

import torch

CFG3_FS = 16000.0
CFG3_SAMPLES = 2048
CFG3_MAX_DELAY = 0.05


def cfg3_frames(frames: int, mics: int = 32, seed: int = 3000, device="cuda", chunk: int = 512,
                n: int = CFG3_SAMPLES, fs: float = CFG3_FS) -> torch.Tensor:
    """[frames, mics, n] float32: src = 0.5*N(0,1) + 0.5*speech-like (Hann-windowed 800/1150/2900 Hz
    formants of signal_processing.py:40-48); channel m = src delayed by d ~ U{0..39} samples
    + 0.3*N(0,1)."""
    dev = torch.device(device)
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    pad = 64
    t = torch.arange(n + pad, device=dev, dtype=torch.float64) / fs
    win = torch.hann_window(n + pad, periodic=False, device=dev, dtype=torch.float64)
    speech = ((torch.sin(2 * math.pi * 800 * t) + 0.8 * torch.sin(2 * math.pi * 1150 * t + math.pi / 4)
               + 0.5 * torch.sin(2 * math.pi * 2900 * t + math.pi / 2)) * win).float()
    out = torch.empty((frames, mics, n), dtype=torch.float32, device=dev)
    ar = torch.arange(n, device=dev)
    for f0 in range(0, frames, chunk):
        nb = min(chunk, frames - f0)
        src = 0.5 * torch.randn((nb, n + pad), generator=g, device=dev) + 0.5 * speech
        d = torch.randint(0, 40, (nb, mics), generator=g, device=dev)
        idx = (40 - d)[:, :, None] + ar[None, None, :]
        ch = src[:, None, :].expand(nb, mics, n + pad).gather(2, idx)
        ch += 0.3 * torch.randn((nb, mics, n), generator=g, device=dev)
        out[f0:f0 + nb] = ch
    return out
