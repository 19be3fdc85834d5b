#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k2_colpass_fwd|k_whiten' --launch-skip 14 -c 4 -o gpurun_out/s14_cfg5_fwd python tools/bench_configs.py cfg5 --small > gpurun_out/s14_ncu.log 2>&1
ls -la gpurun_out/*.ncu-rep
