#!/bin/bash
# GPU session 2: second-generation convolution engine -- parity, A/B against the first engine, chunk sweep, launch list
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/s2_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/s2_tests.log
tail -4 gpurun_out/s2_tests.log
for f in 0 1; do
  PAL_FFT2=$f timeout 300 python tools/bench_configs.py cfg2 cfg5 > gpurun_out/s2_configs_fft2_$f.jsonl 2> gpurun_out/s2_configs_fft2_$f.err
  echo "== PAL_FFT2=$f"; cut -c1-330 gpurun_out/s2_configs_fft2_$f.jsonl
done
for mb in 24 48 96 192; do
  echo "== chunk $mb MB"
  PAL_CONV_CHUNK_MB=$mb timeout 300 python tools/bench_configs.py cfg2 cfg5 2>/dev/null | grep gcc_phat_tdoa | cut -c1-330 | tee -a gpurun_out/s2_chunks_$mb.jsonl
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/s2_launches_cfg.csv python tools/bench_configs.py cfg2 cfg5 > gpurun_out/s2_ncu.log 2>&1
echo done
