#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -s > gpurun_out/s5_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/s5_tests.log
tail -4 gpurun_out/s5_tests.log; grep -i "batched solve" gpurun_out/s5_tests.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/s5_bench.json 2> gpurun_out/s5_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/s5_bench.json'))
s=d['scenes']
print("cfg3", d['value'], "e2e", d['e2e']['value'], "parity", d['parity']['ok'])
print("scenes", s['value'], s['split_ms'], s['parity']['ok'], s['ends_in'])
PY
