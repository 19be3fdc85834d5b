#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/s4_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/s4_tests.log
tail -4 gpurun_out/s4_tests.log
timeout 600 python tools/bench_configs.py cfg5 cfg4 2>gpurun_out/s4_cfg.err | cut -c1-420 | tee gpurun_out/s4_configs.jsonl
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/s4_bench.json 2> gpurun_out/s4_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/s4_bench.json'))
s=d['scenes']
print("cfg3", d['value'], "e2e", d['e2e']['value'], "parity", d['parity']['ok'], d['parity']['lag_mismatches'])
print("scenes", s['value'], s['split_ms'], s['parity'])
PY
