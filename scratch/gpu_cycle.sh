#!/bin/bash
# usage: scratch/gpu_cycle.sh TAG [notest]   -- run on the GPU box: parity tests, bench line, one ncu --set full capture
TAG=$1
if [ "$2" != "notest" ]; then python -m pytest tests -m gpu -x -q 2>&1 | tail -4; fi
python bench.py --no-cpu --no-e2e > gpurun_out/${TAG}_bench.json 2>gpurun_out/${TAG}_bench.err
python - <<PY
import json
d = json.load(open('gpurun_out/${TAG}_bench.json'))
print('ms_per_step', d['ms_per_step'], 'frac', d['roofline']['frac'])
PY
ncu --set full --clock-control none --import-source on -k regex:mate_step_kernel2 -s 10 -c 1 -o gpurun_out/${TAG}_full -f python bench.py --steps 20 --warmup 3 --no-cpu --no-e2e > gpurun_out/${TAG}_ncu.log 2>&1
tail -1 gpurun_out/${TAG}_ncu.log
