#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -s > gpurun_out/s8_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/s8_tests.log
tail -4 gpurun_out/s8_tests.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/s8_launches_sweep.csv python tools/bench_cfg5.py --scenes-per-gpu 4096 --chunk 4096 --steps 1 --warmup 1 > gpurun_out/s8_ncu.log 2>&1
python tools/ncu_summary.py launches gpurun_out/s8_launches_sweep.csv | head -30
