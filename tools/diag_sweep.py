"""Host-side timeline of one sweep step (where does the host wait?)."""
import sys, time, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pyaudiolocalization_b200 import sweep, scene as _scene, gcc_phat as _g

cfg = sweep.SweepConfig()
n, chunk = 32768, 16384
sw = sweep.SceneSweep(cfg, n, chunk=chunk, parts=int(sys.argv[1]) if len(sys.argv) > 1 else 4)
sets = [sweep.random_shoebox_scenes(n, cfg.mics, 5000 + 1000 * i) for i in range(5)]
for i in range(3):
    sw.step(*sets[i])
torch.cuda.synchronize()
for i in (3, 4):
    src, mic, pl = sets[i]
    t = [time.perf_counter()]
    lab = []
    cur = torch.cuda.current_stream()
    chunks = [(c0, min(c0 + chunk, n)) for c0 in range(0, n, chunk)]
    sw.geom_stream.wait_stream(cur)
    job = sw._prepare(src, mic, pl, *chunks[0]); t.append(time.perf_counter()); lab.append("prepare0")
    for ci, (c0, c1) in enumerate(chunks):
        cur.wait_event(job.done)
        sig = _scene.execute_render(job, plan_cache=sw.cache, grouped_parts=sw.parts); t.append(time.perf_counter()); lab.append(f"render{ci}")
        res = _g.gcc_phat_tdoa_batched(sig, float(cfg.fs), cfg.max_expected_delay); t.append(time.perf_counter()); lab.append(f"gcc{ci}")
        sw.k_all[c0:c1] = res.k_idx
        if ci + 1 < len(chunks):
            job = sw._prepare(src, mic, pl, *chunks[ci + 1]); t.append(time.perf_counter()); lab.append(f"prepare{ci+1}")
    torch.cuda.synchronize(); t.append(time.perf_counter()); lab.append("final sync")
    print("step", i, "total %.1f ms:" % ((t[-1] - t[0]) * 1e3), ", ".join(f"{l} {1e3*(b-a):.1f}" for l, a, b in zip(lab, t[:-1], t[1:])), flush=True)
print("alloc stats: reserved %.1f GB, allocated peak %.1f GB, cudaMalloc retries %d" % (torch.cuda.memory_reserved() / 1e9, torch.cuda.max_memory_allocated() / 1e9, torch.cuda.memory_stats().get("num_alloc_retries", 0)))
