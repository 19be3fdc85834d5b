#!/bin/bash
# source-level (SASS) stall sampling of the fused pair kernel: one full ncu capture, 4096 frames; the CSV pages are exported
# on the box (the report itself is about 40 MB)
mkdir -p gpurun_out
B="bench.py --steps 2 --warmup 1 --no-e2e --no-cpu --no-parity --no-scenes --frames 4096"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_pair4095_tmem' --launch-skip 1 -c 1 -f -o /tmp/s31_pair python $B > gpurun_out/s31_ncu_pair.log 2>&1; echo "ncu pair rc=$?"
ncu -i /tmp/s31_pair.ncu-rep --page source --csv --print-source sass > gpurun_out/s31_pair_sass.csv 2> gpurun_out/s31_exp.err
ncu -i /tmp/s31_pair.ncu-rep --page raw --csv > gpurun_out/s31_pair_raw.csv 2>> gpurun_out/s31_exp.err
ncu -i /tmp/s31_pair.ncu-rep --page source --csv --print-source cuda,sass > gpurun_out/s31_pair_cuda_sass.csv 2>> gpurun_out/s31_exp.err
gzip -9 gpurun_out/s31_pair_cuda_sass.csv
ls -la gpurun_out/s31* /tmp/s31_pair.ncu-rep
