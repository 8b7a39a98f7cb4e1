#!/bin/bash
# round 2, GPU call 3: why is mate_step_kernel3 slower? timing variants + two ncu captures
cd /root/repo
mkdir -p gpurun_out
for spec in k3a_x3 k3c_x3 k3a_x1 k3a_x2 k3c_x2 k3a_x4 k3c_x4 k3b_x4 k3a k3c; do
  tag=$spec
  MATE_B200_LIB=/root/repo/scratch/variants/libmate_$spec.so timeout 300 python bench.py --no-cpu --no-e2e --steps 1000 --warmup 20 > gpurun_out/r2c_$tag.json 2>gpurun_out/r2c_$tag.err
  python - <<PY
import json
try:
    d = json.load(open('gpurun_out/r2c_$tag.json')); print('$spec', round(d['ms_per_step'],5), round(d['roofline']['frac'],4), d['clocks']['sm_mhz'])
except Exception as e: print('$spec', 'FAILED', e, open('gpurun_out/r2c_$tag.err').read()[-800:])
PY
done
for v in k3a k3c; do
MATE_B200_LIB=/root/repo/scratch/variants/libmate_$v.so timeout 600 ncu --set full --clock-control none --import-source on -k regex:mate_step_kernel3 -s 10 -c 1 -o gpurun_out/r2c_${v}_full -f python bench.py --steps 20 --warmup 3 --no-cpu --no-e2e > gpurun_out/r2c_${v}_ncu.log 2>&1
tail -2 gpurun_out/r2c_${v}_ncu.log
done
