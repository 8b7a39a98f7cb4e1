#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
echo "== parity (auto tile), then odd tile sizes"
timeout 900 python -m pytest tests/test_cuda_parity.py -m gpu -x -q 2>&1 | tail -3
MATE_B200_TILE=28 timeout 900 python -m pytest tests/test_cuda_parity.py -m gpu -x -q 2>&1 | tail -3
MATE_B200_TILE=5 timeout 900 python -m pytest tests/test_cuda_parity.py -m gpu -x -q -k "oracle or prepared" 2>&1 | tail -3
echo "== timeline tile 28"
MATE_B200_LIB=/root/repo/scratch/variants/libmate_t28.so timeout 300 python scratch/timeline.py gpurun_out/r2i_timeline28.npy 2>&1 | tail -15
run() { # name lib extra-args
  MATE_B200_LIB=/root/repo/scratch/variants/libmate_$2.so timeout 300 python bench.py --no-cpu --no-e2e ${@:3} > gpurun_out/r2i_$1.json 2>gpurun_out/r2i_$1.err
  python - <<PY
import json
try:
    d = json.load(open('gpurun_out/r2i_$1.json')); print('$1', round(d['ms_per_step'],5), round(d['roofline']['frac'],4), d['clocks']['sm_mhz'])
except Exception as e: print('$1', 'FAILED', e, open('gpurun_out/r2i_$1.err').read()[-800:])
PY
}
for t in 32 30 28 27 26 24 22 20 16; do MATE_B200_TILE=$t run tile$t base --steps 1000 --warmup 20; done
run auto base --steps 1000 --warmup 20
for i in 1 2 3; do run auto_drv$i base --steps 20 --warmup 5; done
