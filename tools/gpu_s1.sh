#!/bin/bash
# GPU session 1 of round 2: new parity tests, bench line with parity + scenes, secondary configs
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv > gpurun_out/s1_gpu.txt 2>&1
nproc >> gpurun_out/s1_gpu.txt
timeout 1200 python -m pytest tests -m gpu -x -q -s > gpurun_out/s1_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/s1_tests.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/s1_bench.json 2> gpurun_out/s1_bench.err; echo "bench rc=$?" >> gpurun_out/s1_bench.err
timeout 300 python tools/bench_configs.py cfg2 cfg5 > gpurun_out/s1_configs.jsonl 2> gpurun_out/s1_configs.err
tail -3 gpurun_out/s1_tests.log; cat gpurun_out/s1_bench.json | cut -c1-1500; tail -3 gpurun_out/s1_bench.err
