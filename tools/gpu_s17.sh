#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/s17_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/s17_tests.log
tail -3 gpurun_out/s17_tests.log
for v in 1 0; do
echo "== PAL_SMEM_CONV=$v"
PAL_SMEM_CONV=$v timeout 300 python tools/bench_configs.py cfg5 2>gpurun_out/s17_$v.err | grep '"gcc_phat_tdoa"' | cut -c1-300
PAL_SMEM_CONV=$v timeout 300 python tools/bench_short_frames.py 2>/dev/null | cut -c1-300
done
timeout 600 python tools/bench_cfg5.py --scenes-per-gpu 32768 --chunk 16384 --steps 3 > gpurun_out/s17_cfg5.json 2>gpurun_out/s17_cfg5.err
python - <<PY
import json
d=json.load(open('gpurun_out/s17_cfg5.json')); print("sweep", "%.0f scenes/s" % d['value'], "%.1f ms" % d['ms_per_step'], d['split_ms'], d['gpu_launches'])
PY
