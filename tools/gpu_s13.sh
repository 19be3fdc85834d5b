#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/s13_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/s13_tests.log
tail -3 gpurun_out/s13_tests.log
timeout 600 python tools/bench_configs.py cfg2 cfg5 cfg4 2>gpurun_out/s13_cfg.err | cut -c1-330 | tee gpurun_out/s13_configs.jsonl
timeout 600 python tools/bench_cfg5.py --scenes-per-gpu 32768 --chunk 16384 --steps 3 > gpurun_out/s13_cfg5.json 2>gpurun_out/s13_cfg5.err
python - <<PY
import json
d=json.load(open('gpurun_out/s13_cfg5.json')); print("sweep", "%.0f scenes/s" % d['value'], "%.1f ms" % d['ms_per_step'], d['split_ms'], d['gpu_launches'])
PY
