#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_transfer_group|k_render_group' --launch-skip 8 -c 4 -o gpurun_out/s16_render python tools/bench_cfg5.py --scenes-per-gpu 8192 --chunk 8192 --steps 1 --warmup 2 --parts 1 > gpurun_out/s16_ncu.log 2>&1
ls -la gpurun_out/s16_render.ncu-rep
