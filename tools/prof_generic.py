"""Small driver for ncu captures of the arbitrary-length GCC-PHAT path: python tools/prof_generic.py [cfg5|cfg2] [B]"""
import sys

import torch

sys.path.insert(0, ".")
import pyaudiolocalization_b200 as pal

which = sys.argv[1] if len(sys.argv) > 1 else "cfg5"
if which == "cfg5":
    B, M, N, fs = 1024, 8, 4000, 16000.0
else:
    B, M, N, fs = 32, 4, 44100, 44100.0
if len(sys.argv) > 2:
    B = int(sys.argv[2])
g = torch.Generator(device="cuda"); g.manual_seed(1)
fr = torch.randn((B, M, N), generator=g, device="cuda")
fr[:, 1:] = 0.5 * fr[:, :1] + 0.5 * fr[:, 1:]
for _ in range(2):
    r = pal.gcc_phat_tdoa_batched(fr, fs, 0.05)
torch.cuda.synchronize()
print("ok", float(((r.flags & 8) != 0).float().mean()))
