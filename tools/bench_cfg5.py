"""BASELINE.json configs[4] ("throughput sweep"): random shoebox scenes (8 mics, 0.25 s @ 16 kHz chirp,
max_reflections = 3) rendered by the batched stage 1 and fed to the batched GCC-PHAT / TDOA stage, scenes
sharded over the GPUs of one box, one NCCL all-gather of the per-scene lag vectors per step.

    python tools/bench_cfg5.py [--scenes-per-gpu 16384] [--chunk 4096] [--steps 2]
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/bench_cfg5.py

Prints one JSON line (rank 0).  Scene geometry is generated on the host before the timed region;
everything from the image sources to the gathered lag indices is inside it.
"""
import argparse
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import pyaudiolocalization_b200 as pal  # noqa: E402
from pyaudiolocalization_b200 import main as pmain, shard  # noqa: E402

MATS = {"air": {"absorption": 0.01, "freq": 1e-6}, "wood": {"absorption": 0.05, "freq": 1e-5},
        "metal": {"absorption": 0.1, "freq": 2e-5}, "glass": {"absorption": 0.07, "freq": 1.5e-5}}
FS, DUR, FREQ, MICS, ORDER, MED = 16000, 0.25, 500, 8, 3, 0.05
ROOM_MATERIALS = ["wood", "metal", "glass", "wood", "wood", "metal"]      # planes x=0, x=lx, y=0, y=ly, z=0, z=lz


def shoebox(lx, ly, lz):
    m = ["wood", "metal", "glass", "wood", "wood", "metal"]
    pl = [[1, 0, 0, 0], [1, 0, 0, -lx], [0, 1, 0, 0], [0, 1, 0, -ly], [0, 0, 1, 0], [0, 0, 1, -lz]]
    return [{"plane": p, "material": mm} for p, mm in zip(pl, m)]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scenes-per-gpu", type=int, default=16384)
    ap.add_argument("--chunk", type=int, default=16384)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=1)
    ap.add_argument("--no-plan-cache", action="store_true", help="rebuild the renderer's per-length tables for every bucket")
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
    s_n = args.scenes_per_gpu
    from pyaudiolocalization_b200 import scene
    from pyaudiolocalization_b200.signal_processing import generate_signal

    def scene_set(seed):
        rng = np.random.default_rng(seed)
        dims = rng.uniform([3, 3, 2.5], [10, 8, 4], size=(s_n, 3))
        mics = 0.3 + rng.uniform(size=(s_n, MICS, 3)) * (dims[:, None, :] - 0.6)
        srcs = 0.3 + rng.uniform(size=(s_n, 3)) * (dims - 0.6)
        # rooms in the array form of scene.image_sources_batched: coefficients [S, 6, 4] + the six material names
        pl = np.zeros((s_n, 6, 4))
        pl[:, 0, 0] = pl[:, 1, 0] = pl[:, 2, 1] = pl[:, 3, 1] = pl[:, 4, 2] = pl[:, 5, 2] = 1.0
        pl[:, 1, 3], pl[:, 3, 3], pl[:, 5, 3] = -dims[:, 0], -dims[:, 1], -dims[:, 2]
        return srcs, mics, pl

    # a DIFFERENT random scene set for every step (as in a real sweep), generated before the timed region
    sets = [scene_set(5000 + rank + 1000 * i) for i in range(args.warmup + args.steps)]
    base = torch.as_tensor(generate_signal("chirp", FS, DUR, FREQ).astype(np.float32)).to(dev)
    cache = None if args.no_plan_cache else scene.RenderPlanCache()
    P = MICS * (MICS - 1) // 2
    k_all = torch.empty((s_n, P, 1), dtype=torch.int32, device=dev)
    gathered = torch.empty((world, s_n, P, 1), dtype=torch.int32, device=dev) if world > 1 else None
    step_no = [0]

    def step():
        srcs, mics, rooms = sets[step_no[0] % len(sets)]
        step_no[0] += 1
        for c0 in range(0, s_n, args.chunk):
            c1 = min(c0 + args.chunk, s_n)
            sig = pmain.simulate_scenes_batched(srcs[c0:c1], mics[c0:c1], FS, 343.62, DUR, "chirp", FREQ,
                                                (rooms[c0:c1], ROOM_MATERIALS), MATS, ORDER, 0.01, base_signal=base,
                                                plan_cache=cache)
            res = pal.gcc_phat_tdoa_batched(sig, float(FS), MED)
            k_all[c0:c1] = res.k_idx
        if world > 1:
            dist.all_gather_into_tensor(gathered, k_all)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    barrier()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0 = pal.launch_count()
    t0.record()
    for _ in range(args.steps):
        step()
    t1.record()
    barrier()
    ms = t0.elapsed_time(t1) / args.steps
    launches = (pal.launch_count() - l0) // args.steps
    if world > 1:
        tt = torch.tensor([ms], device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms = float(tt.item())
    if rank == 0:
        td = shard.tdoa_seconds_from_indices(k_all[:4], int(DUR * FS), FS)
        line = {"metric": "scenes_per_s (render + GCC-PHAT TDOA)", "value": world * s_n / (ms * 1e-3), "unit": "scenes/s",
                "pair_corr_per_s": world * s_n * P / (ms * 1e-3), "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "dtype": "f32 (f64 geometry / phases)",
                "data": "synthetic",
                "config": {"workload": "cfg5: random shoebox scenes, 8 mics, 0.25 s @ 16 kHz chirp 500 Hz, max_reflections=3, "
                                       "rendered then GCC-PHAT TDOA (28 pairs, n = 7999), max_expected_delay=0.05 s",
                           "scenes_per_gpu": s_n, "chunk_scenes": args.chunk,
                           "scene_sets": "a different random set per step, generated before the timed region",
                           "render_plan_cache": None if cache is None else
                           {"hits": cache.hits, "misses": cache.misses, "plans": len(cache.plans), "bytes": cache.bytes},
                           "parallelism": f"scenes sharded over {world} GPU(s); one NCCL all-gather of lag indices per step"},
                "gpu_launches": int(launches), "sample_tdoa_s": [float(x) for x in td[0, :4, 0]]}
        out_stream.write(json.dumps(line) + "\n")
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
