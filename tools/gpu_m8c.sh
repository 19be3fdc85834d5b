#!/bin/bash
# final build on 8 GPUs: the driver's own command line (default flags)
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/m8c_bench.json 2> gpurun_out/m8c.err
echo "rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/m8c_bench.json'))
s=d.get('scenes') or {}
print('cfg3 %.5g ms %.2f e2e %.5g parity %s rows %s' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['parity']['ok'], d['parity']['rows']))
print('scenes %.5g ms %.2f parity %s' % (s.get('value',0), s.get('ms_per_step',0), (s.get('parity') or {}).get('ok')))
PY
tail -3 gpurun_out/m8c.err
