"""Raw concurrent host->device copy bandwidth of this box: the floor of every host-buffer ("e2e") number.

    python tools/h2d_bw.py                                   # one GPU
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/h2d_bw.py

Every rank copies a pinned 1 GiB buffer to its GPU `reps` times, all ranks starting together; rank 0 prints one JSON line
with the per-rank rates, the aggregate (bytes of all ranks over the slowest rank's time) and the same with the copy split
over 2 and 4 streams.  Also times device->host for the result size of the cfg3 step (130 MB).
"""
import json
import os
import sys

import torch
import torch.distributed as dist


def main():
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    nbytes = 1 << 30
    reps = 8
    host = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
    host.fill_(1)
    devbuf = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    out = {}
    for n_streams in (1, 2, 4):
        streams = [torch.cuda.Stream(dev) for _ in range(n_streams)]
        part = nbytes // n_streams
        for _ in range(2):      # warm-up
            devbuf.copy_(host, non_blocking=True)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        for s in streams:
            s.wait_stream(torch.cuda.current_stream())
        for _ in range(reps):
            for i, s in enumerate(streams):
                with torch.cuda.stream(s):
                    devbuf[i * part:(i + 1) * part].copy_(host[i * part:(i + 1) * part], non_blocking=True)
        for s in streams:
            torch.cuda.current_stream().wait_stream(s)
        t1.record()
        torch.cuda.synchronize()
        ms = t0.elapsed_time(t1)
        rate = reps * nbytes / (ms * 1e-3) / 1e9
        rates = [rate]
        if world > 1:
            t = torch.tensor([rate], device=dev, dtype=torch.float64)
            allr = [torch.zeros_like(t) for _ in range(world)]
            dist.all_gather(allr, t)
            rates = [float(x.item()) for x in allr]
        out[f"h2d_streams_{n_streams}"] = {"per_rank_GBps": [round(r, 2) for r in rates],
                                           "aggregate_GBps": round(world * min(rates), 2)}
    # device -> host, 130 MB (k_idx + gmax + tdoa of one cfg3 step)
    small = 130023424
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if world > 1:
        dist.barrier()
    t0.record()
    for _ in range(reps):
        host[:small].copy_(devbuf[:small], non_blocking=True)
    t1.record()
    torch.cuda.synchronize()
    out["d2h_130MB_GBps_rank0"] = round(reps * small / (t0.elapsed_time(t1) * 1e-3) / 1e9, 2)
    if rank == 0:
        out.update({"n_gpus": world, "bytes_per_copy": nbytes, "reps": reps, "host_cores": os.cpu_count()})
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
