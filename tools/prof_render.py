"""Small driver for ncu launch lists of the renderer: python tools/prof_render.py [cfg4|cfg5]"""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
which = sys.argv[1] if len(sys.argv) > 1 else "cfg4"
sys.argv = ["x"]
import tools.bench_configs as bc
from pyaudiolocalization_b200 import main as pmain

if which == "cfg4":
    mics = np.random.default_rng(0).uniform([1, 1, 0.5], [5, 4, 2.5], size=(64, 3))
    srcs = np.random.default_rng(1).uniform([0.5, 0.5, 0.3], [5.5, 4.5, 2.7], size=(16, 3))
    args = (srcs, mics, 48000, 343.62, 1.0, "chirp", 1000, bc.shoebox(6, 5, 3), bc.MATS, 6, 0.01)
else:
    rng = np.random.default_rng(5000)
    rooms, micl, srcl = [], [], []
    for _ in range(2048):
        dims = rng.uniform([3, 3, 2.5], [10, 8, 4])
        rooms.append(bc.shoebox(*dims))
        micl.append(rng.uniform([0.3, 0.3, 0.3], dims - 0.3, size=(8, 3)))
        srcl.append(rng.uniform([0.3, 0.3, 0.3], dims - 0.3))
    args = (np.array(srcl), np.array(micl), 16000, 343.62, 0.25, "chirp", 500, rooms, bc.MATS, 3, 0.01)
for _ in range(2):
    out = pmain.simulate_scenes_batched(*args)
torch.cuda.synchronize()
print("ok", tuple(out.shape))
