#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k2_conv_smem' --launch-skip 6 -c 3 -o gpurun_out/s20_smemconv python tools/bench_configs.py cfg5 --small > gpurun_out/s20_ncu.log 2>&1
ls -la gpurun_out/s20_smemconv.ncu-rep
