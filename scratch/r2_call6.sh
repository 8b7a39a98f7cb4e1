#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
for spec in k3a_x3 k3a_x2p2 k3a_x2p2s0 k3a_x2p6 k3a_x2p2tma k3a_x2p6tma k3a_x2 k3a_x2tma k3a_p6 k3a_p4 k3a_p6@MATE_B200_KERNEL=2 k3a_p4@MATE_B200_KERNEL=2; do
  n=${spec%%@*}; envs=""
  if [[ "$spec" == *@* ]]; then envs=$(echo "${spec#*@}" | tr ',' ' '); fi
  tag=$(echo "$spec" | tr '@=,' '___')
  env $envs MATE_B200_LIB=/root/repo/scratch/variants/libmate_$n.so timeout 300 python bench.py --no-cpu --no-e2e --steps 1000 --warmup 20 > gpurun_out/r2f_$tag.json 2>gpurun_out/r2f_$tag.err
  python - <<PY
import json
try:
    d = json.load(open('gpurun_out/r2f_$tag.json')); print('$spec', round(d['ms_per_step'],5), round(d['roofline']['frac'],4), d['clocks']['sm_mhz'])
except Exception as e: print('$spec', 'FAILED', e, open('gpurun_out/r2f_$tag.err').read()[-800:])
PY
done
