#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -s > gpurun_out/s9_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/s9_tests.log
tail -4 gpurun_out/s9_tests.log
for ch in 16384; do
timeout 600 python tools/bench_cfg5.py --scenes-per-gpu 32768 --chunk $ch --steps 2 > gpurun_out/s9_cfg5_$ch.json 2>gpurun_out/s9_cfg5_$ch.err
done
python - <<'PY'
import json
for ch in (16384,):
    try:
        d=json.load(open('gpurun_out/s9_cfg5_%d.json'%ch)); print(ch, "%.0f scenes/s" % d['value'], "%.1f ms" % d['ms_per_step'], d['split_ms'], d['gpu_launches'])
    except Exception as e: print(ch, "failed", e)
PY
tail -3 gpurun_out/s9_cfg5_16384.err
