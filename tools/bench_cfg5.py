"""BASELINE.json configs[4] ("throughput sweep") on its own: random shoebox scenes (8 mics, 0.25 s @ 16 kHz chirp,
max_reflections = 3) rendered by the batched stage 1 and fed to the batched GCC-PHAT / TDOA stage, scenes sharded over
the GPUs of one box, one NCCL all-gather of the per-scene lag vectors per step.  bench.py runs the same driver
(`pyaudiolocalization_b200.sweep.SceneSweep`) as its `scenes` block; this tool exists for size / chunk sweeps.

    python tools/bench_cfg5.py [--scenes-per-gpu 32768] [--chunk 16384] [--steps 2]
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/bench_cfg5.py
"""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import pyaudiolocalization_b200 as pal  # noqa: E402
from pyaudiolocalization_b200 import sweep  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scenes-per-gpu", type=int, default=32768)
    ap.add_argument("--chunk", type=int, default=16384)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--solve", action="store_true", help="end in source positions (pal_solve_positions)")
    ap.add_argument("--keep", type=int, default=0)
    ap.add_argument("--parts", type=int, default=4, help="streams the grouped renderer spreads its buckets over")
    ap.add_argument("--warmup", type=int, default=3)
    args = ap.parse_args()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    out_stream = os.fdopen(real_stdout, "w", buffering=1)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    cfg = sweep.SweepConfig()
    s_n = args.scenes_per_gpu
    sw = sweep.SceneSweep(cfg, s_n, chunk=args.chunk, device=dev, parts=args.parts, solve=args.solve, keep_signals=args.keep)
    sets = [sweep.random_shoebox_scenes(s_n, cfg.mics, 5000 + rank + 1000 * i) for i in range(args.warmup + args.steps)]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(args.warmup):
        sw.step(*sets[i])
        if i == 0:
            lo, hi = sw.n_seen
            sw.warm_plans(lo - (hi - lo) // 8, hi + (hi - lo) // 8)
    barrier()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0 = pal.launch_count()
    t0.record()
    for i in range(args.steps):
        sw.step(*sets[args.warmup + i])
    t1.record()
    barrier()
    ms = t0.elapsed_time(t1) / args.steps
    launches = (pal.launch_count() - l0) // args.steps
    sw.timed = True
    sw.step(*sets[-1])
    torch.cuda.synchronize()
    if world > 1:
        tt = torch.tensor([ms], device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms = float(tt.item())
    if rank == 0:
        line = {"metric": "scenes_per_s (render + GCC-PHAT TDOA)", "value": world * s_n / (ms * 1e-3), "unit": "scenes/s",
                "pair_corr_per_s": world * s_n * cfg.pairs / (ms * 1e-3), "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
                "dtype": "f32 (f64 geometry / phases)", "data": "synthetic",
                "split_ms": {"render": sw.render_ms, "gcc_phat": sw.gcc_ms},
                "config": {"workload": "cfg5: random shoebox scenes, 8 mics, 0.25 s @ 16 kHz chirp 500 Hz, max_reflections=3, "
                                       "rendered then GCC-PHAT TDOA (28 pairs, n = 7999), max_expected_delay=0.05 s",
                           "scenes_per_gpu": s_n, "chunk_scenes": sw.chunk,
                           "render_plan_cache": {"hits": sw.cache.hits, "misses": sw.cache.misses, "plans": len(sw.cache.plans)}},
                "gpu_launches": int(launches)}
        out_stream.write(json.dumps(line) + "\n")
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
