import sys, numpy as np, torch
sys.path.insert(0, ".")
sys.argv = ["x"]
import tools.bench_configs as bc
from pyaudiolocalization_b200 import main as pmain
which = "cfg4"
rng = np.random.default_rng(1)
mics = np.random.default_rng(0).uniform([1, 1, 0.5], [5, 4, 2.5], size=(64, 3))
srcs = rng.uniform([0.5, 0.5, 0.3], [5.5, 4.5, 2.7], size=(16, 3))
planes = bc.shoebox(6, 5, 3)
for _ in range(2):
    out = pmain.simulate_scenes_batched(srcs, mics, 48000, 343.62, 1.0, "chirp", 1000, planes, bc.MATS, 6, 0.01)
torch.cuda.synchronize()
print("ok", out.shape)
