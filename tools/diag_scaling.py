"""Diagnostic: how the fused pair kernel's per-frame time depends on the batch size, on workspace
chunking and on how long the GPU has been under load (clock behaviour).  Prints one line each."""
import subprocess
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
import pyaudiolocalization_b200 as pal
from pyaudiolocalization_b200 import _lib

FS, MED, MICS, NS = 16000.0, 0.05, 32, 2048
P = MICS * (MICS - 1) // 2


def smi():
    try:
        return subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,clocks.mem,power.draw,temperature.gpu,"
                               "clocks_event_reasons.sw_power_cap,clocks_event_reasons.hw_slowdown,"
                               "clocks_event_reasons.sw_thermal_slowdown", "--format=csv,noheader", "-i", "0"],
                              capture_output=True, text=True).stdout.strip()
    except OSError:
        return "n/a"


def run(B, chunk=None, reps=5):
    fr = pal.synth.cfg3_frames(B, MICS, seed=3000)
    full_bytes, _ = pal.gcc_phat.workspace_bytes(B, MICS, NS, P)
    if chunk:   # spectra for `chunk` frames only (the list region stays sized for B)
        full_bytes -= (B - chunk) * (MICS * 65 * 32 * 8)
    ws = torch.empty(full_bytes + 256, dtype=torch.uint8, device="cuda")
    ks, ke = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ks.record(); ke.record()
    torch.cuda.synchronize()

    def step():
        return pal.gcc_phat_tdoa_batched(fr, FS, MED, workspace=ws)

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(reps):
        step()
    t1.record()
    torch.cuda.synchronize()
    total = t0.elapsed_time(t1) / reps
    clk = smi()
    out = [f"B={B} chunk={chunk} step_ms={total:.3f} frames/s={B / total * 1e3:.0f}"]
    if not chunk:
        for stage in (1, 2, 3):
            _lib.profile_hook(stage, ks, ke)
            tt = []
            for _ in range(3):
                step()
                torch.cuda.synchronize()
                tt.append(ks.elapsed_time(ke))
            out.append(f"stage{stage}_ms={np.mean(tt):.3f}")
        _lib.profile_hook(0)
    out.append(f"smi[{clk}]")
    print(" ".join(out), flush=True)
    del fr, ws
    torch.cuda.empty_cache()


def sustained(B=2048, seconds=3.0):
    """per-step time of back-to-back steps while the GPU stays loaded"""
    fr = pal.synth.cfg3_frames(B, MICS, seed=3000)
    full, _ = pal.gcc_phat.workspace_bytes(B, MICS, NS, P)
    ws = torch.empty(full + 256, dtype=torch.uint8, device="cuda")
    for _ in range(3):
        pal.gcc_phat_tdoa_batched(fr, FS, MED, workspace=ws)
    torch.cuda.synchronize()
    time.sleep(1.0)
    evs = []
    n = 0
    t_start = time.perf_counter()
    e = torch.cuda.Event(enable_timing=True); e.record(); evs.append(e)
    while time.perf_counter() - t_start < seconds:
        for _ in range(10):
            pal.gcc_phat_tdoa_batched(fr, FS, MED, workspace=ws)
            e = torch.cuda.Event(enable_timing=True); e.record(); evs.append(e)
        torch.cuda.synchronize()
        n += 10
        if n % 50 == 0:
            print("   ", smi(), flush=True)
    ms = [evs[i].elapsed_time(evs[i + 1]) for i in range(len(evs) - 1)]
    ms = np.array(ms)
    k = max(1, len(ms) // 8)
    print(f"sustained B={B}: steps={len(ms)} first8th={ms[:k].mean():.3f} ms last8th={ms[-k:].mean():.3f} ms "
          f"min={ms.min():.3f} max={ms.max():.3f}", flush=True)


if __name__ == "__main__":
    print("idle", smi(), flush=True)
    for B in (1024, 2048, 4096, 8192, 16384):
        run(B)
    run(16384, chunk=4096)
    run(16384, chunk=1024)
    sustained()
