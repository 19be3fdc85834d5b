"""Diagnostic: where the end-to-end (host buffer) call spends its time, for several chunk sizes."""
import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
import pyaudiolocalization_b200 as pal

B, M, N = 16384, 32, 2048
fr = pal.synth.cfg3_frames(B, M, seed=3000)
host = torch.empty((B, M, N), dtype=torch.float32, pin_memory=True)
host.copy_(fr); torch.cuda.synchronize()
del fr; torch.cuda.empty_cache()
# copy alone
d = torch.empty((1024, M, N), dtype=torch.float32, device="cuda")
torch.cuda.synchronize(); t = time.perf_counter()
for f0 in range(0, B, 1024):
    d.copy_(host[f0:f0 + 1024], non_blocking=True)
torch.cuda.synchronize(); print("H2D alone, 1024-frame chunks: %.1f ms" % ((time.perf_counter() - t) * 1e3), flush=True)
del d
for chunk in (256, 512, 1024, 2048):
    for rep in range(3):
        torch.cuda.synchronize(); t = time.perf_counter()
        r = pal.gcc_phat.gcc_phat_tdoa_from_host(host, 16000.0, 0.05, chunk_frames=chunk)
        torch.cuda.synchronize(); dt = (time.perf_counter() - t) * 1e3
    print(f"chunk {chunk}: {dt:.1f} ms (last of 3)", flush=True)
