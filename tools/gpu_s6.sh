#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -s > gpurun_out/s6_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/s6_tests.log
tail -4 gpurun_out/s6_tests.log; grep -i "band-limited" gpurun_out/s6_tests.log
for ch in 16384 32768; do
timeout 600 python tools/bench_cfg5.py --scenes-per-gpu 32768 --chunk $ch --steps 2 2>gpurun_out/s6_cfg5_$ch.err | cut -c1-300 | tee gpurun_out/s6_cfg5_$ch.json
done
