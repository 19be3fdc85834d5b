#!/bin/bash
# GPU session 3: reduced window pick + occupancy variants of the convolution engine + ncu of its kernels
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/s3_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/s3_tests.log
tail -4 gpurun_out/s3_tests.log
run() { # name, env...
  name=$1; shift
  echo "== $name"
  env "$@" timeout 300 python tools/bench_configs.py cfg2 cfg5 2>gpurun_out/s3_$name.err | grep '"gcc_phat_tdoa"' | cut -c1-330 | tee gpurun_out/s3_$name.jsonl
}
run mb2
run mb2_fullpick PAL_FAST_PICK=0
run mb3 PAL_B200_LIB=$PWD/build/libpal_mb3.so
run mb4 PAL_B200_LIB=$PWD/build/libpal_mb4.so
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k2_|k_win_pick' --launch-skip 12 -c 14 -o gpurun_out/s3_cfg5_full python tools/bench_configs.py cfg5 --small > gpurun_out/s3_ncu_cfg5.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k2_|k_win_pick' --launch-skip 6 -c 7 -o gpurun_out/s3_cfg2_full python tools/bench_configs.py cfg2 > gpurun_out/s3_ncu_cfg2.log 2>&1
ls -la gpurun_out/*.ncu-rep
echo done
