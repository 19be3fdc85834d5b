#!/bin/bash
# 2-GPU session: sharded-vs-single parity on real GPUs, bench at N = 2 (gather / no gather / NCCL CTA caps), H2D floor
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/m2_gpus.txt
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q -s > gpurun_out/m2_multi_test.log 2>&1; echo "multi test rc=$?" | tee -a gpurun_out/m2_multi_test.log
tail -3 gpurun_out/m2_multi_test.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 900 $TR bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/m2_bench_2gpu.json 2> gpurun_out/m2_bench_2gpu.err; echo "bench2 rc=$?"
timeout 600 $TR bench.py --gpus 2 --steps 10 --warmup 3 --no-e2e --no-scenes --no-parity --no-gather > gpurun_out/m2_bench_2gpu_nogather.json 2> gpurun_out/m2_nog.err
NCCL_MAX_CTAS=2 timeout 600 $TR bench.py --gpus 2 --steps 10 --warmup 3 --no-e2e --no-scenes --no-parity > gpurun_out/m2_bench_2gpu_ctas2.json 2> gpurun_out/m2_c2.err
NCCL_MAX_CTAS=1 timeout 600 $TR bench.py --gpus 2 --steps 10 --warmup 3 --no-e2e --no-scenes --no-parity > gpurun_out/m2_bench_2gpu_ctas1.json 2> gpurun_out/m2_c1.err
timeout 600 $TR bench.py --gpus 2 --steps 10 --warmup 3 --no-e2e --no-scenes --no-parity --scaling strong > gpurun_out/m2_bench_2gpu_strong.json 2> gpurun_out/m2_st.err
timeout 300 python tools/h2d_bw.py > gpurun_out/m2_h2d_1.json 2>/dev/null
timeout 300 $TR tools/h2d_bw.py > gpurun_out/m2_h2d_2.json 2>/dev/null
for f in gpurun_out/m2_bench_2gpu*.json; do echo $f; python -c "
import json,sys
d=json.load(open('$f'))
print(' value %.4g ms/step %.2f scaling %s gather %s' % (d['value'], d['ms_per_step'], d['scaling'], d['config'].get('gather')))
if d.get('e2e'): print(' e2e %.4g' % d['e2e']['value'])
if d.get('parity'): print(' parity', d['parity']['ok'], d['parity']['rows'], d['parity']['lag_mismatches'])
if d.get('scenes'): print(' scenes %.4g' % d['scenes']['value'], d['scenes']['parity']['ok'])
"; done
cat gpurun_out/m2_h2d_1.json gpurun_out/m2_h2d_2.json
