#!/bin/bash
# 2-GPU session b: SMs reserved for the concurrent all-gather (0 / 2 / 4), with and without an NCCL CTA cap
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_solver.py tests/test_gpu_multi.py -m gpu -x -q -s > gpurun_out/m2b_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/m2b_tests.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512"
B="bench.py --gpus 2 --steps 10 --warmup 3 --no-e2e --no-scenes --no-parity"
for r in 0 2 4; do
  timeout 600 $TR $B --reserve-sms $r > gpurun_out/m2b_reserve${r}.json 2> gpurun_out/m2b_r$r.err
  NCCL_MAX_CTAS=2 timeout 600 $TR $B --reserve-sms $r > gpurun_out/m2b_reserve${r}_ctas2.json 2> gpurun_out/m2b_r${r}c.err
done
timeout 600 $TR $B --no-gather --reserve-sms 0 > gpurun_out/m2b_nogather.json 2> gpurun_out/m2b_ng.err
timeout 600 python bench.py --steps 10 --warmup 3 --no-e2e --no-scenes --no-parity --no-cpu > gpurun_out/m2b_1gpu.json 2> gpurun_out/m2b_1.err
for f in gpurun_out/m2b_*.json; do python -c "
import json
d=json.load(open('$f'))
print('$f'.split('/')[-1], ' value %.5g  ms/step %.2f  reserved %s gather %s' % (d['value'], d['ms_per_step'], d['config'].get('reserved_sms'), d['config'].get('gather')))
"; done
