#!/usr/bin/env python
"""bench.py -- GCC-PHAT mic-pair correlations/s of the B200 path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--frames F]

A "step" is one pass of the hot path (forward transforms -> fused cross-spectrum + PHAT +
inverse DFT + peak pick -> float64 re-evaluation of flagged rows) over one batch of the cfg3
workload of BASELINE.json / SURVEY.md section 8d: 16384 frames x 32 mics x 2048 samples
(496 pairs per frame, n = 4095), PHAT TDOA with max_expected_delay = 0.05 s at fs = 16 kHz,
PER GPU (weak scaling).  Inputs are resident in HBM before the timed region (4.3 GB per
GPU, far larger than L2, so no L2 flush is needed between iterations).  With N > 1 the ranks
shard frames with no data-path collective and one NCCL all-gather of the per-frame lag
indices at the end of every step (inside the timed region).

The JSON line carries: value (device-resident throughput, all GPUs), e2e (same metric through
the public host-buffer API: pinned host frames -> H2D -> kernels -> D2H of lag indices, every
step), roofline (fused pair kernel, algorithmic bytes of SURVEY section 8d over its CUDA-event
time, against MEASURED_PEAKS.json), cpu_baseline (the oracle port of the reference, timed on
this box's host cores on a bounded sample).  `--impl reference` times that CPU port alone.
"""
import argparse
import json
import multiprocessing as mp
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FS = 16000.0
MED = 0.05
MICS = 32
NS = 2048
PAIRS = MICS * (MICS - 1) // 2
# 20*M*N + 16*P*N + 16*P bytes per frame for the three-kernel decomposition
ALG_BYTES_PER_FRAME = 20 * MICS * NS + 16 * PAIRS * NS + 16 * PAIRS
COMPULSORY_BYTES_PER_FRAME = 4 * MICS * NS + 16 * PAIRS
METRIC = "gcc_phat_pair_correlations_per_s"
UNIT = "pair-corr/s"


def numpy_cfg3_frames(frames, mics, seed):
    """Host twin of pyaudiolocalization_b200.synth.cfg3_frames (same recipe, numpy RNG)."""
    rng = np.random.default_rng(seed)
    n, pad = NS, 64
    t = np.arange(n + pad) / FS
    speech = (np.sin(2 * np.pi * 800 * t) + 0.8 * np.sin(2 * np.pi * 1150 * t + np.pi / 4)
              + 0.5 * np.sin(2 * np.pi * 2900 * t + np.pi / 2)) * np.hanning(n + pad)
    out = np.empty((frames, mics, n), np.float32)
    for f in range(frames):
        src = 0.5 * rng.standard_normal(n + pad) + 0.5 * speech
        d = rng.integers(0, 40, size=mics)
        for m in range(mics):
            out[f, m] = src[40 - d[m]:40 - d[m] + n] + 0.3 * rng.standard_normal(n)
    return out


# ----------------------------------------------------------------------------- CPU reference arm
def _cpu_worker(frame):
    from oracle import pal_oracle as O
    sig = [frame[m].astype(np.float64) for m in range(frame.shape[0])]
    tds, _, cm = O.pair_loop(sig, FS, max_expected_delay=MED)
    return len(tds)


def cpu_reference_rate(n_frames, mics, cores, steps=1, warmup=0):
    """pair-corr/s of the oracle port (same numpy/scipy calls as the reference's
    get_time_delays_phat loop, main.py:202-228) with one process per host core."""
    frames = numpy_cfg3_frames(n_frames, mics, 777)
    ctx = mp.get_context("fork")
    times = []
    with ctx.Pool(cores) as pool:
        pool.map(_cpu_worker, [frames[0][:4]] * cores)           # start-up / imports excluded
        for it in range(warmup + steps):
            t0 = time.perf_counter()
            done = sum(pool.map(_cpu_worker, list(frames), chunksize=1))
            dt = time.perf_counter() - t0
            if it >= warmup:
                times.append((done, dt))
    done = sum(d for d, _ in times)
    dt = sum(t for _, t in times)
    return done / dt, dt / len(times), done // len(times)


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    n_frames = args.ref_frames or 16 * cores
    rate, sec_per_step, per_step = cpu_reference_rate(n_frames, MICS, cores, steps=args.steps, warmup=min(args.warmup, 1))
    line = {
        "impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": min(args.warmup, 1), "ms_per_step": sec_per_step * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "cfg3: 32 mics x 2048-sample frames, 496 pairs, PHAT TDOA, max_expected_delay=0.05 s, fs=16 kHz",
                   "frames_per_step": n_frames, "pairs_per_frame": PAIRS},
        "cpu_baseline": {"value": rate, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{n_frames} frames x {PAIRS} pairs per step ({per_step} pair-corr), oracle port of "
                                   "utils.get_time_delays_phat looped as main.py:202-228, one process per core"},
        "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ----------------------------------------------------------------------------- clocks sampler
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.rows.append(ln.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            p = [x.strip() for x in r.split(",")]
            if len(p) < 6:
                continue
            try:
                sm.append(float(p[0]))
                mx.append(float(p[1]))
            except ValueError:
                continue
            for nme, v in zip(names, p[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(nme)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------------------------------- GPU arm
def run_gpu_arm(args):
    import torch
    import torch.distributed as dist

    import pyaudiolocalization_b200 as pal
    from pyaudiolocalization_b200 import _lib, synth

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the product path has no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    frames_n = args.frames
    frames = synth.cfg3_frames(frames_n, MICS, seed=3000 + rank, device=dev)
    full, _ = pal.gcc_phat.workspace_bytes(frames_n, MICS, NS, PAIRS)
    ws = torch.empty(full + 256, dtype=torch.uint8, device=dev)
    pairs_dev = torch.from_numpy(pal.all_pairs(MICS)).to(dev)
    def new_out():
        return pal.TdoaBatch(torch.empty((frames_n, PAIRS, 1), dtype=torch.int32, device=dev),
                             torch.empty((frames_n, PAIRS), dtype=torch.int32, device=dev),
                             torch.empty((frames_n, PAIRS), dtype=torch.float32, device=dev),
                             torch.empty((frames_n, PAIRS), dtype=torch.float32, device=dev),
                             torch.empty((frames_n, PAIRS), dtype=torch.int32, device=dev), None, NS, FS)
    # Results are double-buffered so that the NCCL all-gather of step i (asynchronous, on NCCL's own
    # stream) overlaps the kernels of step i+1: a rank never idles inside a step waiting for a slower
    # peer; every gather has completed before the timed region ends.
    outs = [new_out(), new_out()] if world > 1 else [new_out()]
    out = outs[0]
    gathered = [torch.empty((world, frames_n, PAIRS, 1), dtype=torch.int32, device=dev) for _ in outs] if world > 1 else None
    works = [None, None]
    step_no = [0]

    def step():
        s = step_no[0] % len(outs)
        step_no[0] += 1
        if works[s] is not None:
            works[s].wait()
        pal.gcc_phat_tdoa_batched(frames, FS, MED, workspace=ws, out=outs[s], pairs_dev=pairs_dev)
        if world > 1 and not args.no_gather:
            works[s] = dist.all_gather_into_tensor(gathered[s], outs[s].k_idx, async_op=True)

    def drain():
        for s in range(len(works)):
            if works[s] is not None:
                works[s].wait()
                works[s] = None

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # per-kernel timing of the dominant kernel (fused pair kernel) through the library's hook
    ks, ke = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ks.record(); ke.record()
    torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    drain()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    l0 = pal.launch_count()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    kernel_ms = []
    t0.record()
    for _ in range(args.steps):
        step()
    drain()
    t1.record()
    barrier()
    ms = t0.elapsed_time(t1)
    launches = pal.launch_count() - l0
    # one more pass with the stage hook, outside the headline timing, for the kernel duration
    for stage in (1, 2, 3):
        _lib.profile_hook(stage, ks, ke)
        reps = []
        for _ in range(max(2, min(args.steps, 5))):
            step()
            drain()
            torch.cuda.synchronize()
            reps.append(ks.elapsed_time(ke))
        kernel_ms.append(float(np.mean(reps)))
    _lib.profile_hook(0)
    clocks = sampler.stop() if rank == 0 else None
    if world > 1:
        tt = torch.tensor([ms], device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms = float(tt.item())
    if world > 1 and not args.no_gather:   # sharded + gathered lags of the last step: this rank's slice must be its own result
        assert torch.equal(gathered[(step_no[0] - 1) % 2][rank], outs[(step_no[0] - 1) % 2].k_idx)
    flags = out.flags
    refined_frac = float(((flags & 8) != 0).float().mean().item())

    # ---- end-to-end through the public host-buffer API ------------------------------------
    e2e = e2e_pcm16 = None
    if not args.no_e2e:
        host = torch.empty((frames_n, MICS, NS), dtype=torch.float32, pin_memory=True)
        host.copy_(frames)
        torch.cuda.synchronize()
        out_host = pal.gcc_phat.alloc_host_outputs(frames_n, PAIRS)      # pinned result buffers, re-used every step
        r = pal.gcc_phat.gcc_phat_tdoa_from_host(host, FS, MED, chunk_frames=args.e2e_chunk, out_host=out_host)   # warm-up
        assert np.array_equal(r["k_idx"], out.k_idx.cpu().numpy())
        barrier()
        e_steps = max(1, min(args.steps, 3))
        w0 = time.perf_counter()
        for _ in range(e_steps):
            r = pal.gcc_phat.gcc_phat_tdoa_from_host(host, FS, MED, chunk_frames=args.e2e_chunk, out_host=out_host)
        barrier()
        e_sec = (time.perf_counter() - w0) / e_steps
        if world > 1:
            tt = torch.tensor([e_sec], device=dev, dtype=torch.float64)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            e_sec = float(tt.item())
        e2e = {"value": world * frames_n * PAIRS / e_sec, "unit": UNIT, "h2d_bytes_per_step": r["h2d_bytes"],
               "d2h_bytes_per_step": r["d2h_bytes"], "ms_per_step": e_sec * 1e3, "steps": e_steps}
        del host
        # The same call fed with 16-bit PCM host buffers (what capture hardware / WAV files deliver): half the
        # PCIe bytes, converted to float32 x/32768 on the device.  Informational; `e2e` above is the float32 contract.
        q = torch.clamp(torch.round(frames * (32767.0 / float(frames.abs().max()))), -32768, 32767).to(torch.int16)
        host16 = torch.empty((frames_n, MICS, NS), dtype=torch.int16, pin_memory=True)
        host16.copy_(q)
        fq = pal.gcc_phat.pcm16_to_f32(q)
        del q
        want = pal.gcc_phat_tdoa_batched(fq, FS, MED, pairs_dev=pairs_dev).k_idx.cpu().numpy()
        del fq
        r16 = pal.gcc_phat.gcc_phat_tdoa_from_host(host16, FS, MED, chunk_frames=args.e2e_pcm_chunk, out_host=out_host)   # warm-up
        assert np.array_equal(r16["k_idx"], want)
        barrier()
        w0 = time.perf_counter()
        for _ in range(e_steps):
            r16 = pal.gcc_phat.gcc_phat_tdoa_from_host(host16, FS, MED, chunk_frames=args.e2e_pcm_chunk, out_host=out_host)
        barrier()
        p_sec = (time.perf_counter() - w0) / e_steps
        if world > 1:
            tt = torch.tensor([p_sec], device=dev, dtype=torch.float64)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            p_sec = float(tt.item())
        e2e_pcm16 = {"value": world * frames_n * PAIRS / p_sec, "unit": UNIT, "h2d_bytes_per_step": r16["h2d_bytes"],
                     "d2h_bytes_per_step": r16["d2h_bytes"], "ms_per_step": p_sec * 1e3, "steps": e_steps,
                     "input": "int16 PCM host frames (the synthetic frames quantised to 16 bits), float32 on the device"}
        del host16

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except OSError:
        pass
    peak_gbs = float(peaks.get("hbm_gbs", 6650.0))
    traffic, traffic_src = None, None
    try:   # DRAM bytes of the pair kernel from the committed ncu --set full capture, scaled to this launch size
        tj = json.load(open(os.path.join(ROOT, "profiles", "r1e_traffic.json")))
        traffic = float(tj["dram_bytes_per_frame"]) * frames_n
        traffic_src = tj["source"]
    except (OSError, KeyError, ValueError):
        pass
    ms_step = ms / args.steps
    value = world * frames_n * PAIRS / (ms_step * 1e-3)
    achieved = ALG_BYTES_PER_FRAME * frames_n / (kernel_ms[1] * 1e-3) / 1e9
    cores = os.cpu_count() or 1
    cpu = None
    if not args.no_cpu and world == 1:      # the CPU baseline is reported at N = 1 only
        nfr = 48 * cores      # about 10-20 s of work on the box's cores
        rate, sec, per = cpu_reference_rate(nfr, MICS, cores)
        cpu = {"value": rate, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": f"{nfr} cfg3 frames x {PAIRS} pairs ({per} pair-corr, {sec:.1f} s), oracle port of "
                         "utils.get_time_delays_phat looped as main.py:202-228, one process per host core"}
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": "cfg3: 32 mics x 2048-sample frames, 496 pairs, PHAT TDOA, max_expected_delay=0.05 s, fs=16 kHz",
                   "frames_per_gpu": frames_n, "pairs_per_frame": PAIRS, "n_fft": 2 * NS - 1,
                   "l2": "inputs (4.3 GB/GPU at 16384 frames) exceed L2; no flush needed",
                   "parallelism": f"frames sharded over {world} GPU(s); one NCCL all-gather of lag indices per step"
                   if world > 1 else "1 GPU", "frames_per_s": value / PAIRS,
                   "refined_row_fraction": refined_frac},
        "clocks": clocks,
        "e2e": e2e,
        "e2e_pcm16": e2e_pcm16,
        "gpu_launches": int(launches),
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak_gbs, "unit": "GB/s",
                     "frac": achieved / peak_gbs, "traffic": traffic,
                     "traffic_unit": "DRAM bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum)",
                     "traffic_source": traffic_src,
                     "kernel": "k_pair4095_tmem (fused cross-spectrum of the whitened spectra + inverse DFT-4095 + peak pick)",
                     "kernel_ms": kernel_ms[1], "forward_ms": kernel_ms[0], "refine_ms": kernel_ms[2],
                     "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6650 GB/s",
                     "algorithmic_bytes_per_frame": ALG_BYTES_PER_FRAME,
                     "compulsory_bytes_per_frame": COMPULSORY_BYTES_PER_FRAME,
                     "note": "achieved uses SURVEY 8d's three-kernel byte count; the fused kernel keeps R on "
                             "chip, so its real bound is the FP32 pipe, see DESIGN.md"},
        "cpu_baseline": cpu,
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    # The contract is ONE JSON line on stdout: libraries that print there (NCCL's version banner)
    # are sent to stderr for the duration of the run.
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    sys.stdout = os.fdopen(real_stdout, "w", buffering=1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--frames", type=int, default=16384, help="frames per GPU per step (cfg3: 16384)")
    ap.add_argument("--ref-frames", type=int, default=0, help="frames per step of the CPU reference arm")
    ap.add_argument("--e2e-chunk", type=int, default=256)
    ap.add_argument("--e2e-pcm-chunk", type=int, default=1024,
                    help="frames per pipeline chunk of the int16 host path (compute-bound: fewer, larger launches)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-gather", action="store_true", help="diagnostic only: skip the NCCL all-gather (invalid as a result)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_gpu_arm(args)


if __name__ == "__main__":
    main()
