#!/bin/bash
# 8-GPU session: weak + strong scaling of the bench line, raw concurrent H2D floor at 2 / 4 / 8 ranks
mkdir -p gpurun_out
nvidia-smi -L | wc -l > gpurun_out/m8_gpus.txt; nproc >> gpurun_out/m8_gpus.txt
TR8="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521"
TR4="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29522"
TR2="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29523"
timeout 900 $TR8 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/m8_bench_8gpu.json 2> gpurun_out/m8_bench_8gpu.err; echo "bench8 rc=$?"
timeout 600 $TR8 bench.py --gpus 8 --steps 10 --warmup 3 --no-e2e --no-scenes --no-parity --scaling strong > gpurun_out/m8_bench_8gpu_strong.json 2> gpurun_out/m8_s.err
timeout 600 $TR8 bench.py --gpus 8 --steps 10 --warmup 3 --no-e2e --no-scenes --no-parity --no-gather > gpurun_out/m8_bench_8gpu_nogather.json 2> gpurun_out/m8_ng.err
timeout 600 $TR4 bench.py --gpus 4 --steps 10 --warmup 3 --no-scenes --no-parity > gpurun_out/m8_bench_4gpu.json 2> gpurun_out/m8_4.err
timeout 300 $TR2 tools/h2d_bw.py > gpurun_out/m8_h2d_2.json 2>/dev/null
timeout 300 $TR4 tools/h2d_bw.py > gpurun_out/m8_h2d_4.json 2>/dev/null
timeout 300 $TR8 tools/h2d_bw.py > gpurun_out/m8_h2d_8.json 2>/dev/null
timeout 300 python tools/h2d_bw.py > gpurun_out/m8_h2d_1.json 2>/dev/null
for f in gpurun_out/m8_bench_*.json; do python -c "
import json
d=json.load(open('$f'))
print('$f'.split('/')[-1], ' value %.5g  ms/step %.2f  %s gather %s' % (d['value'], d['ms_per_step'], d['scaling'], d['config'].get('gather')))
if d.get('e2e'): print('   e2e %.4g  (%.1f ms, %.1f GB/s per GPU)' % (d['e2e']['value'], d['e2e']['ms_per_step'], d['e2e'].get('h2d_GBps_per_gpu', 0)))
if d.get('parity'): print('   parity', d['parity']['ok'], d['parity']['rows'], d['parity']['lag_mismatches'])
if d.get('scenes'): print('   scenes %.4g' % d['scenes']['value'], d['scenes']['parity']['ok'], d['scenes']['ms_per_step'])
"; done
cat gpurun_out/m8_h2d_*.json
