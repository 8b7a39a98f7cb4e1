#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
for spec in k3p2000 k3p2500 k3p3000 k3p3500 k3p4000 k3p3000c k3p3000x2 k3p2500x2 k3a; do
  n=${spec%%@*}; envs=""
  if [[ "$spec" == *@* ]]; then envs=$(echo "${spec#*@}" | tr ',' ' '); fi
  tag=$(echo "$spec" | tr '@=,' '___')
  env $envs MATE_B200_LIB=/root/repo/scratch/variants/libmate_$n.so timeout 300 python bench.py --no-cpu --no-e2e --steps 1000 --warmup 20 > gpurun_out/r2g_$tag.json 2>gpurun_out/r2g_$tag.err
  python - <<PY
import json
try:
    d = json.load(open('gpurun_out/r2g_$tag.json')); print('$spec', round(d['ms_per_step'],5), round(d['roofline']['frac'],4), d['clocks']['sm_mhz'])
except Exception as e: print('$spec', 'FAILED', e, open('gpurun_out/r2g_$tag.err').read()[-800:])
PY
done
