#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
timeout 180 python - <<'PY'
# smoke with the dev variant would need the full lib; check parity of the variant on the 4v8-9 cases instead
PY
for spec in k3a_x3 k3e_x3 k3a_x2 k3a_x4 k3a k3a_tma k3e k3a; do
  MATE_B200_LIB=/root/repo/scratch/variants/libmate_$spec.so timeout 300 python bench.py --no-cpu --no-e2e --steps 1000 --warmup 20 > gpurun_out/r2d_$spec.json 2>gpurun_out/r2d_$spec.err
  python - <<PY
import json
try:
    d = json.load(open('gpurun_out/r2d_$spec.json')); print('$spec', round(d['ms_per_step'],5), round(d['roofline']['frac'],4), d['clocks']['sm_mhz'])
except Exception as e: print('$spec', 'FAILED', e, open('gpurun_out/r2d_$spec.err').read()[-800:])
PY
done
MATE_B200_LIB=/root/repo/scratch/variants/libmate_k3a.so timeout 600 python -m pytest tests/test_cuda_parity.py -m gpu -x -q -k "4v8-9" 2>&1 | tail -3
MATE_B200_LIB=/root/repo/scratch/variants/libmate_k3a.so timeout 600 ncu --set full --clock-control none --import-source on -k regex:mate_step_kernel3 -s 10 -c 1 -o gpurun_out/r2d_k3a_full -f python bench.py --steps 20 --warmup 3 --no-cpu --no-e2e > gpurun_out/r2d_k3a_ncu.log 2>&1
tail -1 gpurun_out/r2d_k3a_ncu.log
