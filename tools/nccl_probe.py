"""Probe: NCCL all-gather time for the per-step lag-index payload (run under torchrun)."""
import os, torch, torch.distributed as dist
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
for dtype, name in ((torch.int32, "int32"), (torch.int16, "int16")):
    x = torch.zeros((16384, 496), dtype=dtype, device="cuda")
    out = torch.empty((world, 16384, 496), dtype=dtype, device="cuda")
    for _ in range(3):
        dist.all_gather_into_tensor(out, x)
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        dist.all_gather_into_tensor(out, x)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    if rank == 0:
        print(f"all_gather {name}: {ms:.3f} ms per call, {x.numel()*x.element_size()*(world-1)/ms/1e6:.1f} GB/s in per rank", flush=True)
if rank == 0:
    print("p2p access 0->1:", torch.cuda.can_device_access_peer(0, 1) if world > 1 else None)
dist.destroy_process_group()
