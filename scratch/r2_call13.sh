#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
echo "== smoke + parity (full lib, a+b+c on)"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python -m pytest tests/test_cuda_parity.py -m gpu -x -q 2>&1 | tail -3
run() { # name lib extra-args
  MATE_B200_LIB=/root/repo/scratch/variants/libmate_$2.so timeout 300 python bench.py --no-cpu --no-e2e ${@:3} > gpurun_out/r2l_$1.json 2>gpurun_out/r2l_$1.err
  python - <<PY
import json
try:
    d = json.load(open('gpurun_out/r2l_$1.json')); print('$1', round(d['ms_per_step'],5), round(d['roofline']['frac'],4), d['clocks']['sm_mhz'])
except Exception as e: print('$1', 'FAILED', e, open('gpurun_out/r2l_$1.err').read()[-800:])
PY
}
for r in 1 2; do for v in base a b c abc; do run ${v}_$r $v --steps 1000 --warmup 20; done; done
MATE_B200_LIB=/root/repo/scratch/variants/libmate_abc_tl.so timeout 300 python scratch/timeline.py gpurun_out/r2l_timeline.npy 2>&1 | tail -15
