#!/bin/bash
# forward-kernel variants: in-tree build (64 regs, 4 blocks/SM), 75-register build, HEAD build
for v in "" variants/libB.so variants/libhead.so; do
  PAL_B200_LIB=$v timeout 600 python bench.py --steps 10 --warmup 3 --no-scenes --no-parity > gpurun_out/s26_b.json 2> gpurun_out/s26_b.err
  python - "$v" <<'PY'
import json,sys
d=json.load(open("gpurun_out/s26_b.json"))
print(sys.argv[1] or "in-tree", d["value"], d["ms_per_step"], "fwd GB/s", d["roofline"]["forward_kernel"]["achieved"])
PY
done
