#!/bin/bash
mkdir -p gpurun_out
TR8="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531"
B="bench.py --gpus 8 --steps 10 --warmup 3 --no-e2e --no-scenes --no-parity"
timeout 600 $TR8 $B --reserve-sms 2 > gpurun_out/m8b_reserve2.json 2> gpurun_out/m8b_r2.err
timeout 600 $TR8 $B --reserve-sms 4 > gpurun_out/m8b_reserve4.json 2> gpurun_out/m8b_r4.err
NCCL_MAX_CTAS=4 timeout 600 $TR8 $B --reserve-sms 4 > gpurun_out/m8b_reserve4_ctas4.json 2> gpurun_out/m8b_r4c.err
timeout 600 $TR8 $B --reserve-sms 0 > gpurun_out/m8b_reserve0.json 2> gpurun_out/m8b_r0.err
for f in gpurun_out/m8b_*.json; do python -c "
import json
d=json.load(open('$f'))
print('$f'.split('/')[-1], ' value %.5g  ms/step %.2f  reserved %s' % (d['value'], d['ms_per_step'], d['config'].get('reserved_sms')))
"; done
