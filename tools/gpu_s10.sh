#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/s10_launches_sweep.csv python tools/bench_cfg5.py --scenes-per-gpu 8192 --chunk 8192 --steps 1 --warmup 1 > gpurun_out/s10_ncu.log 2>&1
python tools/ncu_summary.py launches gpurun_out/s10_launches_sweep.csv 2>/dev/null | head -24
