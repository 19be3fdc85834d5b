#!/usr/bin/env python
"""bench.py -- GCC-PHAT mic-pair correlations/s (+ scenes/s) of the B200 path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--frames F] [--scaling weak|strong]

A "step" is one pass of the hot path (forward transforms -> fused cross-spectrum + PHAT + inverse DFT + peak pick
-> float64 re-evaluation of flagged rows) over one batch of the cfg3 workload of BASELINE.json / SURVEY.md 8d:
16384 frames x 32 mics x 2048 samples (496 pairs per frame, n = 4095), PHAT TDOA with max_expected_delay = 0.05 s at
fs = 16 kHz, PER GPU (weak scaling; `--scaling strong` splits 16384 frames over the ranks).  Inputs are resident in
HBM before the timed region (4.3 GB per GPU, far larger than L2, so no L2 flush is needed between iterations).  With
N > 1 the ranks shard frames with no data-path collective and one NCCL all-gather of the per-frame lag indices per step
(inside the timed region), driven by the package's own `shard.ShardedTdoa`.

The JSON line carries
  value        device-resident throughput, all GPUs
  e2e          the same metric through the public host-buffer API (pinned host frames -> H2D -> kernels -> D2H)
  roofline     fused pair kernel: algorithmic bytes of SURVEY 8d over its CUDA-event time, against MEASURED_PEAKS.json
  parity       GPU vs the reference's CPU implementation on the first 64 frames x 496 pairs of every rank's batch
               (lag indices / TDOAs bit-exact, max(corr) within 1e-4) -- computed in THIS run
  scenes       the scenes/s half of the metric: BASELINE cfg5 sweep (random rooms rendered by stage 1, fed to
               stage 2, lag vectors all-gathered), with its own parity verdict and roofline numbers
  cpu_baseline the reference's own CPU implementation (baseline/_ref, unmodified; else the oracle port) timed on this
               box's host cores on a bounded sample.  `--impl reference` times that CPU arm alone.
"""
import argparse
import json
import logging
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
# the pair kernel's own share of those bytes: read spectra once (8MN), write + read R (16PN), write results (16P)
PAIR_KERNEL_ALG_BYTES_PER_FRAME = 8 * MICS * NS + 16 * PAIRS * NS + 16 * PAIRS
METRIC = "gcc_phat_pair_correlations_per_s"
UNIT = "pair-corr/s"
CORR_RTOL = 1e-4          # north_star: correlation values within 1e-4 relative (fp32)
RENDER_ATOL = 1e-5        # north_star: rendered mic signals within 1e-5 absolute
WORKLOAD = "cfg3: 32 mics x 2048-sample frames, 496 pairs, PHAT TDOA, max_expected_delay=0.05 s, fs=16 kHz"


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
_REF = None


def reference_utils():
    """The UNMODIFIED reference's `utils` module from baseline/_ref (tools/install_reference.py), or False."""
    global _REF
    if _REF is None:
        try:
            from baseline import ref_loader
            _REF = ref_loader.load() if ref_loader.available() else False
        except Exception:      # noqa: BLE001
            _REF = False
        logging.getLogger().setLevel(logging.ERROR)     # the reference warns on every fall-back branch
    return _REF


def cpu_kind():
    return "reference" if reference_utils() else "port"


def _cpu_pairs(args):
    """All i<j pairs of one frame [M, N] as the reference's loop main.py:202-228 does them: `get_time_delays_phat`
    with num_peaks=1 and the frame's max(corr).  Returns (tdoa seconds [P], max(corr) [P])."""
    frame, fs, med = args
    sig = [np.ascontiguousarray(frame[m], dtype=np.float64) for m in range(frame.shape[0])]
    U = reference_utils()
    if not U:
        from oracle import pal_oracle as U          # port of the same calls (no reference copy on this box)
    td, gm = [], []
    for i in range(len(sig)):
        for j in range(i + 1, len(sig)):
            t, corr, _ = U.get_time_delays_phat(sig[i], sig[j], fs, num_peaks=1, max_expected_delay=med)
            td.append(t[0])
            gm.append(np.max(corr))
    return np.array(td, np.float64), np.array(gm, np.float64)


def _cpu_scene(args):
    """One cfg5 scene through the reference's stage 1 (main.simulate_signals_with_multipath) -> [M, n] float64."""
    src, mics, planes, mats, cfg = args
    if reference_utils():
        import importlib
        R = importlib.import_module("main")        # baseline/_ref/main.py (stubs installed by ref_loader)
    else:
        from oracle import pal_oracle as R
    out = R.simulate_signals_with_multipath(src, mics, cfg["fs"], cfg["c"], duration=cfg["duration"], signal_type=cfg["signal_type"],
                                            freq=cfg["freq"], reflective_planes=planes, material_properties=mats,
                                            max_reflections=cfg["max_reflections"], absorption_threshold=cfg["thr"])
    return np.array(out)


_POOL = None


def host_pool(cores=None):
    """One process per host core, forked ONCE before this process touches CUDA / NCCL (forking afterwards is not
    safe), kept for the CPU baseline and the in-run parity checks.  Start-up and imports are paid here."""
    global _POOL
    if _POOL is None:
        cores = cores or (os.cpu_count() or 1)
        pool = mp.get_context("fork").Pool(cores)
        warm = numpy_cfg3_frames(1, 4, 1)[0]
        pool.map(_cpu_pairs, [(warm, FS, MED)] * cores)
        _POOL = (pool, cores)
    return _POOL


def cpu_reference_rate(n_frames, mics, steps=1, warmup=0):
    """pair-corr/s of the reference's get_time_delays_phat loop with one process per host core."""
    frames = numpy_cfg3_frames(n_frames, mics, 777)
    pool, cores = host_pool()
    times = []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        done = sum(len(r[0]) for r in pool.map(_cpu_pairs, [(f, FS, MED) for f in frames], chunksize=1))
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append((done, dt))
    done = sum(d for d, _ in times)
    dt = sum(t for _, t in times)
    return done / dt, dt / len(times), done // len(times)


def cpu_sample_text(n_frames, per, sec=None):
    what = ("the UNMODIFIED reference (baseline/_ref/utils.py get_time_delays_phat" if cpu_kind() == "reference"
            else "oracle port of the reference (utils.get_time_delays_phat")
    return (f"{n_frames} cfg3 frames x {PAIRS} pairs ({per} pair-corr" + (f", {sec:.1f} s" if sec else "") + f"), {what} "
            "looped as main.py:202-228), one process per host core")


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    _, cores = host_pool()
    n_frames = args.ref_frames or 16 * cores
    rate, sec_per_step, per_step = cpu_reference_rate(n_frames, MICS, steps=args.steps, warmup=min(args.warmup, 1))
    line = {
        "impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": min(args.warmup, 1), "ms_per_step": sec_per_step * 1e3,
        "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "frames_per_step": n_frames, "pairs_per_frame": PAIRS},
        "cpu_baseline": {"value": rate, "unit": UNIT, "cores": cores, "kind": cpu_kind(),
                         "sample": cpu_sample_text(n_frames, per_step) + " per step"},
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


# ----------------------------------------------------------------------------- parity inside the run
def parity_cfg3(frames, res, world, rank, dist, dev, torch, n_par=64):
    """First `n_par` frames x 496 pairs of THIS rank's batch through the reference's CPU implementation on the host
    pool (SURVEY.md 8d parity subset): TDOA seconds must be bit-identical, max(corr) within 1e-4 relative."""
    n_par = min(n_par, frames.shape[0])
    host = frames[:n_par].cpu().numpy()
    td_gpu = res.tdoa_seconds()[:n_par, :, 0]
    gm_gpu = res.gmax[:n_par].double().cpu().numpy()
    refined = int(((res.flags[:n_par] & 8) != 0).sum().item())
    pool, _ = host_pool()
    t0 = time.perf_counter()
    out = pool.map(_cpu_pairs, [(host[f], FS, MED) for f in range(n_par)], chunksize=1)
    sec = time.perf_counter() - t0
    td_ref = np.stack([o[0] for o in out])
    gm_ref = np.stack([o[1] for o in out])
    mism = int((td_gpu != td_ref).sum())
    den = np.where(np.abs(gm_ref) > 0, np.abs(gm_ref), 1.0)
    err = float(np.max(np.abs(gm_gpu - gm_ref) / den))
    rows = n_par * td_ref.shape[1]
    if world > 1:
        t = torch.tensor([mism, rows, refined], device=dev, dtype=torch.int64)
        dist.all_reduce(t)
        mism, rows, refined = (int(v) for v in t.tolist())
        e = torch.tensor([err], device=dev, dtype=torch.float64)
        dist.all_reduce(e, op=dist.ReduceOp.MAX)
        err = float(e.item())
    return {"rows": rows, "lag_mismatches": mism, "max_rel_corr_err": err, "corr_rtol": CORR_RTOL, "refined_rows": refined,
            "ok": bool(mism == 0 and err <= CORR_RTOL), "against": cpu_kind(), "frames_per_rank": n_par, "cpu_seconds": sec,
            "what": "TDOA seconds bit-identical and max(corr) within corr_rtol, first frames of every rank's batch"}


# ----------------------------------------------------------------------------- scenes/s: the cfg5 sweep
def run_scenes(args, world, rank, dev, dist, torch, peak_gbs):
    """BASELINE configs[4]: random shoebox scenes (8 mics, 0.25 s @ 16 kHz chirp, max_reflections = 3) rendered by
    stage 1 and fed to stage 2, scenes sharded over the ranks, one all-gather of the lag vectors per step."""
    import pyaudiolocalization_b200 as pal
    from pyaudiolocalization_b200 import sweep
    cfg = sweep.SweepConfig()
    s_n = args.scenes
    n_par = min(args.scene_parity, s_n)
    sw = sweep.SceneSweep(cfg, s_n, chunk=args.scene_chunk, device=dev, keep_signals=n_par, solve=not args.no_scene_solve)
    steps, warm = max(1, min(args.steps, 3)), 3      # three warm-up steps: the renderer's per-length plan cache fills up
    sets = [sweep.random_shoebox_scenes(s_n, cfg.mics, 5000 + rank + 1000 * i) for i in range(warm + steps)]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(warm):
        sw.step(*sets[i])
        if i == 0:      # the renderer's per-length plans for the whole range of padded lengths this workload produces
            lo, hi = sw.n_seen
            planned = sw.warm_plans(lo - (hi - lo) // 8, hi + (hi - lo) // 8)
    barrier()
    l0 = pal.launch_count()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    marks = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
    t0.record()
    marks[0].record()
    for i in range(steps):
        sw.step(*sets[warm + i])
        marks[i + 1].record()
    t1.record()
    barrier()
    ms = t0.elapsed_time(t1) / steps
    step_ms = [marks[i].elapsed_time(marks[i + 1]) for i in range(steps)]
    launches = (pal.launch_count() - l0) // steps
    if world > 1:
        tt = torch.tensor([ms], device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms = float(tt.item())
    # split of a step into render / GCC-PHAT (events per chunk; one extra untimed-for-the-headline step)
    sw.timed = True
    sw.step(*sets[-1])
    sw.timed = False
    torch.cuda.synchronize()
    refined = float(((sw.flags & 8) != 0).float().mean().item())
    # ---- parity on the first scenes of the last set: rendered channels vs the reference's stage 1 (<= 1e-5), and the
    # TDOAs of the GPU-rendered float32 channels vs the reference's stage 2 on those same channels (bit-exact)
    par = None
    if n_par > 0:
        src, mic, pl = sets[-1]
        mats = sweep.SWEEP_MATERIALS
        cd = {"fs": cfg.fs, "c": cfg.c, "duration": cfg.duration, "signal_type": cfg.signal_type, "freq": cfg.freq,
              "max_reflections": cfg.max_reflections, "thr": cfg.absorption_threshold}
        sig = sw.signals.cpu().numpy()
        td_gpu = pal.shard.tdoa_seconds_from_indices(sw.k_all[:n_par], cfg.samples, float(cfg.fs))[:, :, 0]
        pool, _ = host_pool()
        c0 = time.perf_counter()
        ren = pool.map(_cpu_scene, [(src[s], mic[s], sweep.planes_as_dicts(pl[s]), mats, cd) for s in range(n_par)], chunksize=1)
        tdr = pool.map(_cpu_pairs, [(sig[s], float(cfg.fs), cfg.max_expected_delay) for s in range(n_par)], chunksize=1)
        sec = time.perf_counter() - c0
        rerr = float(max(np.max(np.abs(sig[s] - ren[s])) for s in range(n_par)))
        mism = int(sum((td_gpu[s] != tdr[s][0]).sum() for s in range(n_par)))
        if world > 1:
            t = torch.tensor([mism], device=dev, dtype=torch.int64)
            dist.all_reduce(t)
            mism = int(t.item())
            e = torch.tensor([rerr], device=dev, dtype=torch.float64)
            dist.all_reduce(e, op=dist.ReduceOp.MAX)
            rerr = float(e.item())
        par = {"scenes_per_rank": n_par, "rows": world * n_par * cfg.pairs, "lag_mismatches": mism,
               "max_abs_render_err": rerr, "render_atol": RENDER_ATOL, "ok": bool(mism == 0 and rerr <= RENDER_ATOL),
               "against": cpu_kind(), "cpu_seconds": sec,
               "what": "rendered channels vs the reference's simulate_signals_with_multipath; TDOAs of the GPU-rendered "
                       "float32 channels vs the reference's get_time_delays_phat on the same channels, bit-identical"}
    n, m, p = cfg.samples, cfg.mics, cfg.pairs
    alg2 = 20 * m * n + 16 * p * n + 16 * p          # SURVEY 8d stage-2 bytes per scene (2 432 448 at cfg5)
    gcc_s = sw.gcc_ms * 1e-3
    return {"metric": "scenes_per_s (render + GCC-PHAT TDOA + gather)", "value": world * s_n / (ms * 1e-3), "unit": "scenes/s",
            "ms_per_step": ms, "step_ms_rank0": step_ms, "steps": steps, "warmup": warm, "scenes_per_gpu": s_n, "chunk_scenes": sw.chunk,
            "pair_corr_per_s": world * s_n * p / (ms * 1e-3),
            "workload": "cfg5: random shoebox rooms, 8 mics, 0.25 s @ 16 kHz chirp 500 Hz, max_reflections=3, rendered then "
                        "GCC-PHAT TDOA (28 pairs, n = 7999), max_expected_delay=0.05 s; a different scene set per step",
            "ends_in": "source positions (batched bounded least squares on the device, pal_solve_positions) + gathered lag vectors"
                       if sw.solve else "gathered lag vectors",
            "split_ms": {"render": sw.render_ms, "gcc_phat": sw.gcc_ms, "position_solve": sw.solve_ms,
                         "note": "one extra step with events around each stage (synchronising per chunk)"},
            "gcc_scenes_per_s_per_gpu": s_n / gcc_s if gcc_s > 0 else None,
            "render_scenes_per_s_per_gpu": s_n / (sw.render_ms * 1e-3) if sw.render_ms > 0 else None,
            "roofline": {"bound": "hbm", "stage": "gcc_phat (Bluestein path)", "algorithmic_bytes_per_scene": alg2,
                         "achieved": alg2 * s_n / gcc_s / 1e9 if gcc_s > 0 else None, "peak": peak_gbs, "unit": "GB/s",
                         "frac": alg2 * s_n / gcc_s / 1e9 / peak_gbs if gcc_s > 0 else None},
            "refined_row_fraction": refined, "gpu_launches": int(launches), "parity": par,
            "render_plan_cache": {"hits": sw.cache.hits, "misses": sw.cache.misses, "plans": len(sw.cache.plans),
                                  "prebuilt_in_warmup": planned,
                                  "note": "plans depend on (padded length N, base signal) only; the range of N seen in the first "
                                          "warm-up step (+- 1/8) is planned before the timed steps"}}


# ----------------------------------------------------------------------------- GPU arm
def run_gpu_arm(args):
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not (args.no_cpu and args.no_parity and (args.no_scenes or args.scene_parity == 0)):
        host_pool(max(1, (os.cpu_count() or 1) // world))       # before CUDA / NCCL are initialised in this process
    import torch
    import torch.distributed as dist

    import pyaudiolocalization_b200 as pal
    from pyaudiolocalization_b200 import _lib, shard, synth

    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the product path has no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    strong = args.scaling == "strong"
    frames_n = max(1, args.frames // world) if strong else args.frames
    frames = synth.cfg3_frames(frames_n, MICS, seed=3000 + rank, device=dev)
    drv = shard.ShardedTdoa(frames_n, MICS, NS, FS, MED, device=dev, gather=not args.no_gather,
                            reserve_sms=None if args.reserve_sms < 0 else args.reserve_sms)

    def step():
        drv.step(frames)

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
    drv.drain()
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
    drv.drain()
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
            drv.drain()
            torch.cuda.synchronize()
            reps.append(ks.elapsed_time(ke))
        kernel_ms.append(float(np.mean(reps)))
    _lib.profile_hook(0)
    clocks = sampler.stop() if rank == 0 else None
    if world > 1:
        tt = torch.tensor([ms], device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms = float(tt.item())
    drv_reserved = drv.reserved
    gathered, out = drv.last()
    if gathered is not None:   # sharded + gathered lags of the last step: this rank's slice must be its own result
        assert torch.equal(gathered[rank], out.k_idx)
    refined_frac = float(((out.flags & 8) != 0).float().mean().item())
    parity = None if args.no_parity else parity_cfg3(frames, out, world, rank, dist, dev, torch, args.parity_frames)

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
               "d2h_bytes_per_step": r["d2h_bytes"], "ms_per_step": e_sec * 1e3, "steps": e_steps,
               "h2d_GBps_per_gpu": r["h2d_bytes"] / e_sec / 1e9}
        del host
        # The same call fed with 16-bit PCM host buffers (what capture hardware / WAV files deliver): half the
        # PCIe bytes, converted to float32 x/32768 on the device.  Informational; `e2e` above is the float32 contract.
        q = torch.clamp(torch.round(frames * (32767.0 / float(frames.abs().max()))), -32768, 32767).to(torch.int16)
        host16 = torch.empty((frames_n, MICS, NS), dtype=torch.int16, pin_memory=True)
        host16.copy_(q)
        fq = pal.gcc_phat.pcm16_to_f32(q)
        del q
        want = pal.gcc_phat_tdoa_batched(fq, FS, MED, pairs_dev=drv.pairs_dev).k_idx.cpu().numpy()
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
        del host16, out_host

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except OSError:
        pass
    peak_gbs = float(peaks.get("hbm_gbs", 6650.0))
    # ---- the scenes/s half of the metric (cfg5 sweep) ----------------------------------------
    scenes = None
    if not args.no_scenes:
        del frames, drv
        _lib.reserve_sms(0)
        torch.cuda.empty_cache()
        scenes = run_scenes(args, world, rank, dev, dist, torch, peak_gbs)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    traffic, traffic_src = None, None
    for name in ("r2_traffic.json", "r1e_traffic.json"):
        try:   # DRAM bytes of the pair kernel from the committed ncu --set full capture, scaled to this launch size
            tj = json.load(open(os.path.join(ROOT, "profiles", name)))
            traffic = float(tj["dram_bytes_per_frame"]) * frames_n
            traffic_src = tj["source"]
            break
        except (OSError, KeyError, ValueError):
            continue
    ms_step = ms / args.steps
    value = world * frames_n * PAIRS / (ms_step * 1e-3)
    achieved = ALG_BYTES_PER_FRAME * frames_n / (kernel_ms[1] * 1e-3) / 1e9
    cpu = None
    if not args.no_cpu and world == 1:      # the CPU baseline is reported at N = 1 only
        _, cores = host_pool()
        nfr = 48 * cores      # about 10-20 s of work on the box's cores
        rate, sec, per = cpu_reference_rate(nfr, MICS)
        cpu = {"value": rate, "unit": UNIT, "cores": cores, "kind": cpu_kind(), "sample": cpu_sample_text(nfr, per, sec)}
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": WORKLOAD, "frames_per_gpu": frames_n, "pairs_per_frame": PAIRS, "n_fft": 2 * NS - 1,
                   "l2": "inputs (4.3 GB/GPU at 16384 frames) exceed L2; no flush needed",
                   "parallelism": f"frames sharded over {world} GPU(s); one NCCL all-gather of lag indices per step"
                   if world > 1 else "1 GPU", "frames_per_s": value / PAIRS,
                   "refined_row_fraction": refined_frac, "gather": not args.no_gather, "reserved_sms": drv_reserved},
        "clocks": clocks,
        "e2e": e2e,
        "e2e_pcm16": e2e_pcm16,
        "gpu_launches": int(launches),
        "parity": parity,
        "scenes": scenes,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak_gbs, "unit": "GB/s",
                     "frac": achieved / peak_gbs, "traffic": traffic,
                     "traffic_unit": "DRAM bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum)",
                     "traffic_source": traffic_src,
                     "kernel": "k_pair4095_tmem (fused cross-spectrum of the whitened spectra + inverse DFT-4095 + peak pick)",
                     "kernel_ms": kernel_ms[1], "forward_ms": kernel_ms[0], "refine_ms": kernel_ms[2],
                     "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6650 GB/s",
                     "algorithmic_bytes_per_frame": ALG_BYTES_PER_FRAME,
                     "compulsory_bytes_per_frame": COMPULSORY_BYTES_PER_FRAME,
                     "pair_kernel_own_share": {"algorithmic_bytes_per_frame": PAIR_KERNEL_ALG_BYTES_PER_FRAME,
                                               "achieved": PAIR_KERNEL_ALG_BYTES_PER_FRAME * frames_n / (kernel_ms[1] * 1e-3) / 1e9,
                                               "frac": PAIR_KERNEL_ALG_BYTES_PER_FRAME * frames_n / (kernel_ms[1] * 1e-3) / 1e9 / peak_gbs},
                     "whole_step": {"achieved": ALG_BYTES_PER_FRAME * frames_n / (ms_step * 1e-3) / 1e9,
                                    "frac": ALG_BYTES_PER_FRAME * frames_n / (ms_step * 1e-3) / 1e9 / peak_gbs},
                     "forward_kernel": {"bytes_per_frame": 12 * MICS * NS, "achieved": 12 * MICS * NS * frames_n / (kernel_ms[0] * 1e-3) / 1e9,
                                        "frac": 12 * MICS * NS * frames_n / (kernel_ms[0] * 1e-3) / 1e9 / peak_gbs,
                                        "note": "k_fwd4095 really is HBM-bound: 4 B read + 8 B written per sample"},
                     "note": "achieved = SURVEY 8d's whole three-kernel byte count over the fused pair kernel's time (the "
                             "survey's convention for a fused build); pair_kernel_own_share and whole_step restate it; the "
                             "fused kernel keeps R on chip, so its real bound is the FP32 pipe, see DESIGN.md"},
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
    ap.add_argument("--frames", type=int, default=16384, help="frames per GPU per step (cfg3: 16384); total with --scaling strong")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--ref-frames", type=int, default=0, help="frames per step of the CPU reference arm")
    ap.add_argument("--e2e-chunk", type=int, default=256)
    ap.add_argument("--e2e-pcm-chunk", type=int, default=1024,
                    help="frames per pipeline chunk of the int16 host path (compute-bound: fewer, larger launches)")
    ap.add_argument("--parity-frames", type=int, default=64, help="frames per rank checked against the reference inside the run")
    ap.add_argument("--scenes", type=int, default=32768, help="cfg5 scenes per GPU per step of the scenes/s block")
    ap.add_argument("--scene-chunk", type=int, default=16384)
    ap.add_argument("--scene-parity", type=int, default=16, help="cfg5 scenes per rank checked against the reference")
    ap.add_argument("--no-scene-solve", action="store_true", help="scenes block: stop at the gathered lag vectors")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--no-scenes", action="store_true")
    ap.add_argument("--reserve-sms", type=int, default=-1, help="SMs left out of the persistent grids for the concurrent all-gather (default 0, see shard.ShardedTdoa)")
    ap.add_argument("--no-gather", action="store_true", help="diagnostic only: skip the NCCL all-gather (invalid as a result)")
    args = ap.parse_args()
    try:
        if args.impl == "reference":
            run_reference_arm(args)
        else:
            run_gpu_arm(args)
    finally:
        if _POOL is not None:
            _POOL[0].terminate()
            _POOL[0].join()


if __name__ == "__main__":
    main()
