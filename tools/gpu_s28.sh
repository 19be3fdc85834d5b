#!/bin/bash
# round-2 ncu evidence for the bench command itself: launch list + one full capture of the dominant kernels
mkdir -p gpurun_out
B="bench.py --steps 2 --warmup 1 --no-e2e --no-cpu --no-parity --scene-parity 0 --scenes 8192 --scene-chunk 8192"
timeout 600 python $B > gpurun_out/s28_plain.json 2> gpurun_out/s28_plain.err; echo "plain rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/s28_launches.csv python $B > gpurun_out/s28_ncu1.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_pair4095_tmem|k_fwd4095' --launch-skip 2 -c 2 -o gpurun_out/s28_cfg3_full python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu --no-parity --no-scenes > gpurun_out/s28_ncu2.log 2>&1
ls -la gpurun_out/s28*
