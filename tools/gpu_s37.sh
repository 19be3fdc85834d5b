#!/bin/bash
# ncu --set full of the two cfg3 kernels of the pair-table build (one launch each, 16384 frames), same command as
# profiles/r2h_kernels_16384.md
mkdir -p gpurun_out
timeout 75 ncu --set full --clock-control none -k regex:'k_pair4095_tmem|k_fwd4095' --launch-skip 2 -c 2 -f -o gpurun_out/s37_cfg3_full python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu --no-parity --no-scenes > gpurun_out/s37_ncu.log 2>&1; echo "ncu rc=$?"
ls -la gpurun_out/s37*
