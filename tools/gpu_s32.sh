#!/bin/bash
# pair-kernel micro-fixes (odd-column exchange layout, pair table in shared memory, deferred bound): parity tests of the
# 4095 path on the in-tree build, then cfg3-only bench lines of the in-tree build and the look-ahead variants
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_gcc_phat.py tests/test_gpu_configs.py -m gpu -x -q > gpurun_out/s32_tests.log 2>&1; echo "tests rc=$?"; tail -2 gpurun_out/s32_tests.log
for v in "" variants/libla2.so variants/libla4.so; do
  PAL_B200_LIB=$v timeout 300 python bench.py --steps 10 --warmup 3 --no-scenes --no-e2e --no-cpu > gpurun_out/s32_b.json 2> gpurun_out/s32_b.err || tail -3 gpurun_out/s32_b.err
  python - "$v" <<'PY'
import json,sys
d=json.loads(open("gpurun_out/s32_b.json").read().strip().splitlines()[-1])
r=d["roofline"]
print(sys.argv[1] or "in-tree", "ms/step", round(d["ms_per_step"],3), "pair", round(r["kernel_ms"],3), "fwd", round(r["forward_ms"],3), "refine", round(r["refine_ms"],3), "parity", d.get("parity",{}).get("ok"), d.get("parity",{}).get("lag_mismatches"))
PY
  cp gpurun_out/s32_b.json gpurun_out/s32_b_$(basename "${v:-intree}" .so).json
done
