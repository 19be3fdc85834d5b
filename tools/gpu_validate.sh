#!/bin/bash
# validation of the in-tree build on a fresh box: GPU test suite, smoke(), default bench
mkdir -p gpurun_out
date +%s > gpurun_out/s35_t0
timeout 500 python -m pytest tests -m gpu -x -q > gpurun_out/s35_tests.log 2>&1; echo "tests rc=$?" | tee -a gpurun_out/s35_tests.log
tail -3 gpurun_out/s35_tests.log
timeout 120 python __graft_entry__.py smoke > gpurun_out/s35_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/s35_smoke.log
timeout 400 python bench.py > gpurun_out/s35_bench.json 2> gpurun_out/s35_bench.err; echo "bench rc=$?"
date +%s > gpurun_out/s35_t1
python - <<'PY'
import json
d=json.loads(open("gpurun_out/s35_bench.json").read().strip().splitlines()[-1])
print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], "parity", d["parity"]["ok"], "scenes", d["scenes"]["value"], d["scenes"]["parity"]["ok"])
print("wall", int(open("gpurun_out/s35_t1").read())-int(open("gpurun_out/s35_t0").read()))
PY
