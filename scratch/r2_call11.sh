#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
echo "== parity (auto tile), then odd tile sizes"
timeout 900 python -m pytest tests/test_cuda_parity.py -m gpu -x -q 2>&1 | tail -3
MATE_B200_TILE=5 timeout 900 python -m pytest tests/test_cuda_parity.py -m gpu -x -q -k "oracle or prepared" 2>&1 | tail -3
echo "== timeline"
MATE_B200_LIB=/root/repo/scratch/variants/libmate_pool_tl.so timeout 300 python scratch/timeline.py gpurun_out/r2j_timeline.npy 2>&1 | tail -15
run() { # name lib extra-args
  MATE_B200_LIB=/root/repo/scratch/variants/libmate_$2.so timeout 300 python bench.py --no-cpu --no-e2e ${@:3} > gpurun_out/r2j_$1.json 2>gpurun_out/r2j_$1.err
  python - <<PY
import json
try:
    d = json.load(open('gpurun_out/r2j_$1.json')); print('$1', round(d['ms_per_step'],5), round(d['roofline']['frac'],4), d['clocks']['sm_mhz'])
except Exception as e: print('$1', 'FAILED', e, open('gpurun_out/r2j_$1.err').read()[-800:])
PY
}
run pool pool --steps 1000 --warmup 20
for t in 32 30 24; do MATE_B200_TILE=$t run pool_tile$t pool --steps 1000 --warmup 20; done
for i in 1 2 3; do run pool_drv$i pool --steps 20 --warmup 5; done
echo "== all gpu tests"
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
