#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/s12_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/s12_tests.log
tail -3 gpurun_out/s12_tests.log
python tools/diag_sweep.py 4 2>&1 | tail -3
python tools/diag_sweep.py 1 2>&1 | tail -3
for parts in 1 4; do
timeout 600 python tools/bench_cfg5.py --scenes-per-gpu 32768 --chunk 16384 --steps 3 --parts $parts > gpurun_out/s12_cfg5_p$parts.json 2>gpurun_out/s12_cfg5_p$parts.err
python - <<PY
import json
try:
    d=json.load(open('gpurun_out/s12_cfg5_p$parts.json')); print("parts $parts", "%.0f scenes/s" % d['value'], "%.1f ms" % d['ms_per_step'], d['split_ms'], d['gpu_launches'], d['config']['render_plan_cache'])
except Exception as e: print("parts $parts failed", e)
PY
done
