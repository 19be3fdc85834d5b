import torch, numpy as np, sys
sys.path.insert(0,'.')
import pyaudiolocalization_b200 as pal
fr = pal.synth.cfg3_frames(2048, 32, seed=3000)
for eps in (2e-6, 5e-7):
    r = pal.gcc_phat_tdoa_batched(fr, 16000.0, 0.05, tie_eps=eps, refine=False)
    f = r.flags.cpu().numpy()
    n = f.size
    print("eps", eps, {name: float(((f & bit) != 0).sum())/n for name, bit in [("tie",1),("chain",2),("plateau",4),("fallback",16)]}, "any", float(((f&7)!=0).sum())/n)
# fp32 error scale: compare fast corr to fp64 exact kernel corr on 64 rows
r = pal.gcc_phat_tdoa_batched(fr[:4], 16000.0, 0.05, return_corr=True)
c32 = r.corr.cpu().numpy().astype(np.float64)
from oracle import pal_oracle as O
frh = fr[:4].cpu().numpy().astype(np.float64)
pairs = pal.all_pairs(32)
mx = 0
for f_ in range(2):
    for p in range(0, 496, 7):
        i, j = pairs[p]
        c = O.phat_correlation(frh[f_, i], frh[f_, j])
        mx = max(mx, np.abs(c32[f_, p] - c).max())
print("max abs corr err fp32 vs oracle", mx)
