#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/s24_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/s24_tests.log
tail -3 gpurun_out/s24_tests.log
timeout 600 python tools/bench_configs.py cfg2 cfg5 cfg4 img filter 2>gpurun_out/s24_cfg.err > gpurun_out/s24_configs.jsonl; cut -c1-300 gpurun_out/s24_configs.jsonl
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/s24_bench.json 2> gpurun_out/s24_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/s24_bench.json'))
s=d['scenes']
print("cfg3", d['value'], d['ms_per_step'], "e2e", d['e2e']['value'], "parity", d['parity']['ok'], d['parity']['lag_mismatches'], d['parity']['rows'])
print("scenes", s['value'], s['ms_per_step'], s['split_ms'], s['parity']['ok'], s['gpu_launches'])
print("cpu", d['cpu_baseline']['value'], d['cpu_baseline']['kind'], d['cpu_baseline']['cores'])
PY
for s in 1 2; do timeout 500 python tools/soak_parity.py --cases 500 --seed $s > gpurun_out/s24_soak$s.log 2>gpurun_out/s24_soak$s.err; head -c 900 gpurun_out/s24_soak$s.log | head -4; tail -2 gpurun_out/s24_soak$s.err; done
