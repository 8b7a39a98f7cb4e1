#!/bin/bash
# one round of scratch/ab.sh
for spec in "$@"; do
  n=${spec%%@*}; envs=""
  if [[ "$spec" == *@* ]]; then envs=$(echo "${spec#*@}" | tr ',' ' '); fi
  tag=$(echo "$spec" | tr '@=,' '___')
  env $envs MATE_B200_LIB=/root/repo/scratch/variants/libmate_$n.so python bench.py --no-cpu --no-e2e --steps 1000 --warmup 20 ${AB_ARGS} > gpurun_out/ab_$tag.json 2>gpurun_out/ab_$tag.err
  python - <<PY
import json
try:
    d = json.load(open('gpurun_out/ab_$tag.json')); print('$spec', round(d['ms_per_step'],5), round(d['roofline']['frac'],4), d['clocks']['sm_mhz'])
except Exception as e: print('$spec', 'FAILED', e, open('gpurun_out/ab_$tag.err').read()[-500:])
PY
done
