#!/bin/bash
# quick pair-kernel timing on the GPU box: usage tools/quick_bench.sh <tag> [frames...]
tag=$1; shift
for f in "$@"; do
  python bench.py --frames $f --steps 5 --warmup 3 --no-e2e --no-cpu > gpurun_out/bench_${tag}_$f.json 2>gpurun_out/bench_${tag}_$f.err || tail -5 gpurun_out/bench_${tag}_$f.err
done
cat gpurun_out/bench_${tag}_*.json | python -c "
import sys, json
for ln in sys.stdin:
    d=json.loads(ln); r=d['roofline']; print(d['config']['frames_per_gpu'], 'ms/step', round(d['ms_per_step'],3), 'frames/s', round(d['config']['frames_per_s']), 'pair_ms', round(r['kernel_ms'],3), 'fwd', round(r['forward_ms'],3), 'ref', round(r['refine_ms'],3), 'frac', round(r['frac'],3), 'refined', round(d['config']['refined_row_fraction'],5))
"
