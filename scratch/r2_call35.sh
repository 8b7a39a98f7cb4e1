#!/bin/bash
cd /root/repo
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 2>gpurun_out/r2s_n2.err | tail -1 > gpurun_out/r2s_n2.json
python - <<PY
import json
d=json.load(open('gpurun_out/r2s_n2.json'))
print(d['n_gpus'], d['value'], d['ms_per_step'], d['roofline']['frac'], d['e2e'], [(c['workload'], round(c['frac'],3), round(c['env_steps_per_s']/1e6)) for c in d['configs']], d['clocks'])
PY
tail -3 gpurun_out/r2s_n2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 5 --warmup 2 2>/dev/null | tail -1 | cut -c1-300
