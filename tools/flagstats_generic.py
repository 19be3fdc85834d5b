"""Which flags send rows of the arbitrary-length path to the float64 sweep (cfg5-shaped synthetic input)."""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
import pyaudiolocalization_b200 as pal
from pyaudiolocalization_b200.signal_processing import generate_signal

B, M, N, fs, med = 2048, 8, 4000, 16000.0, 0.05
g = torch.Generator(device="cuda"); g.manual_seed(5000)
base = torch.as_tensor(generate_signal("chirp", fs, N / fs, 500.0)[:N], dtype=torch.float32, device="cuda")
d = torch.randint(0, 64, (B, M), generator=g, device="cuda")
idx = torch.arange(N, device="cuda")[None, None, :] - d[:, :, None]
fr = torch.where(idx >= 0, base[idx.clamp(min=0)], torch.zeros((), device="cuda"))
fr = fr + 0.05 * torch.randn((B, M, N), generator=g, device="cuda")
for eps in (1e-6, 3e-7):
    r = pal.gcc_phat_tdoa_batched(fr, fs, med, tie_eps=eps)
    f = r.flags.cpu().numpy()
    n = f.size
    print("eps", eps, {name: round(float(((f & bit) != 0).sum()) / n, 5) for name, bit in
                       [("tie", 1), ("chain", 2), ("plateau", 4), ("refined", 8), ("fallback", 16), ("alt_thr", 32)]}, flush=True)
r = pal.gcc_phat_tdoa_batched(fr, fs, med, return_corr=True)
f = r.flags.cpu().numpy(); pk = r.peak.cpu().numpy(); gm = r.gmax.cpu().numpy(); k = r.k_idx.cpu().numpy()[..., 0]
c = r.corr.cpu().numpy()
rows = np.argwhere((f & 8) != 0)[:8]
dd = d.cpu().numpy()
pairs = pal.all_pairs(M)
for fi, p in rows:
    row = c[fi, p]
    i, j = pairs[p]
    top = np.argsort(row)[-4:][::-1]
    print("frame", fi, "pair", (i, j), "delays", dd[fi, i], dd[fi, j], "k", k[fi, p], "peak", pk[fi, p], "mean|c|", np.abs(row).mean(),
          "top", [(int(t), float(row[t])) for t in top])
c0, W, D = N - 1, 800, 16
for fi, p in rows[:5]:
    row = c[fi, p].astype(np.float64)
    kb = k[fi, p]
    lo, hi = c0 - W - D, c0 + W + D
    seg = row[lo:hi + 1]
    near = [(int(lo + t), float(seg[t] - row[kb])) for t in np.argsort(seg)[-4:][::-1]]
    print("row", fi, p, "k_best", kb, "h", row[kb], "near", near, "mean", np.abs(row).mean())
