#!/bin/bash
mkdir -p gpurun_out
echo "== default (minblocks 4)"
timeout 300 python tools/bench_configs.py cfg2 cfg5 2>gpurun_out/s3b.err | cut -c1-400 | tee gpurun_out/s3b_configs.jsonl
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k2_|k_win_pick' --launch-skip 12 -c 9 -o gpurun_out/s3_cfg5_full python tools/bench_configs.py cfg5 --small > gpurun_out/s3_ncu_cfg5.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k2_|k_win_pick' --launch-skip 6 -c 7 -o gpurun_out/s3_cfg2_full python tools/bench_configs.py cfg2 > gpurun_out/s3_ncu_cfg2.log 2>&1
ls -la gpurun_out/*.ncu-rep
