#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
run() { # name lib extra-args
  MATE_B200_LIB=/root/repo/scratch/variants/libmate_$2.so timeout 300 python bench.py --no-cpu --no-e2e ${@:3} > gpurun_out/r2k_$1.json 2>gpurun_out/r2k_$1.err
  python - <<PY
import json
try:
    d = json.load(open('gpurun_out/r2k_$1.json')); print('$1', round(d['ms_per_step'],5), round(d['roofline']['frac'],4), d['clocks']['sm_mhz'])
except Exception as e: print('$1', 'FAILED', e, open('gpurun_out/r2k_$1.err').read()[-800:])
PY
}
MATE_B200_LIB=/root/repo/scratch/variants/libmate_s14p2.so timeout 600 python -m pytest tests/test_cuda_parity.py -m gpu -x -q -k "4v8-9 or oracle" 2>&1 | tail -3
for v in s14p0 s14p2 s13p3 s14p4; do run $v $v --steps 1000 --warmup 20; done
