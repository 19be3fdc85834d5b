#!/bin/bash
# pair-kernel variants (pair table in shared memory; start-up stagger; look-ahead depth): cfg3-only bench lines of variants/*.so
mkdir -p gpurun_out
for v in variants/*.so; do
  PAL_B200_LIB=$v timeout 300 python bench.py --steps 8 --warmup 3 --no-scenes --no-e2e --no-cpu > gpurun_out/s34_b.json 2> gpurun_out/s34_b.err || tail -3 gpurun_out/s34_b.err
  python - "$v" <<'PY' | tee -a gpurun_out/s34_summary.txt
import json,sys
d=json.loads(open("gpurun_out/s34_b.json").read().strip().splitlines()[-1])
r=d["roofline"]
print(sys.argv[1], "ms/step", round(d["ms_per_step"],3), "pair", round(r["kernel_ms"],3), "fwd", round(r["forward_ms"],3), "refine", round(r["refine_ms"],3), "parity", d.get("parity",{}).get("ok"), d.get("parity",{}).get("lag_mismatches"))
PY
done
