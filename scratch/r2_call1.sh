#!/bin/bash
# round 2, GPU call 1: timing experiments on the round-1 kernel + sanitizer runs
cd /root/repo
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv,noheader
echo "== variants (ms/step, 1000 steps steady state)"
for v in 0 4 6 37 10 49 2 1 0; do
  MATE_B200_LIB=/root/repo/scratch/variants/libmate_x$v.so python bench.py --no-cpu --no-e2e --steps 1000 --warmup 20 > gpurun_out/r2a_x$v.json 2>gpurun_out/r2a_x$v.err
  python - <<PY
import json
try:
    d = json.load(open('gpurun_out/r2a_x$v.json')); print('x$v', round(d['ms_per_step'],5), round(d['roofline']['frac'],4), d['clocks']['sm_mhz'])
except Exception as e: print('x$v', 'FAILED', e, open('gpurun_out/r2a_x$v.err').read()[-500:])
PY
done
echo "== short runs (driver's command) vs warm"
for i in 1 2 3; do python bench.py --steps 20 --warmup 5 --no-cpu --no-e2e | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('20/5', d['ms_per_step'], d['clocks'])"; done
python bench.py --steps 20 --warmup 300 --no-cpu --no-e2e | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('20/300', d['ms_per_step'], d['clocks'])"
python bench.py --steps 200 --warmup 5 --no-cpu --no-e2e | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('200/5', d['ms_per_step'], d['clocks'])"
echo "== write bandwidth"
python scratch/write_bw.py
echo "== sanitizer"
timeout 600 compute-sanitizer --tool memcheck --log-file gpurun_out/r2a_memcheck_smoke.log python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2a_memcheck_smoke.out 2>&1; echo "memcheck smoke rc=$?"; tail -3 gpurun_out/r2a_memcheck_smoke.log
timeout 900 compute-sanitizer --tool racecheck --log-file gpurun_out/r2a_racecheck_smoke.log python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2a_racecheck_smoke.out 2>&1; echo "racecheck smoke rc=$?"; tail -3 gpurun_out/r2a_racecheck_smoke.log
timeout 900 compute-sanitizer --tool racecheck --log-file gpurun_out/r2a_racecheck_autoreset.log python -m pytest tests/test_cuda_parity.py -q -x -m gpu -k "prepared_resets and 4v8" > gpurun_out/r2a_racecheck_autoreset.out 2>&1; echo "racecheck autoreset rc=$?"; tail -3 gpurun_out/r2a_racecheck_autoreset.log; tail -2 gpurun_out/r2a_racecheck_autoreset.out
timeout 900 compute-sanitizer --tool memcheck --log-file gpurun_out/r2a_memcheck_autoreset.log python -m pytest tests/test_cuda_parity.py -q -x -m gpu -k "prepared_resets and 4v8" > gpurun_out/r2a_memcheck_autoreset.out 2>&1; echo "memcheck autoreset rc=$?"; tail -3 gpurun_out/r2a_memcheck_autoreset.log; tail -2 gpurun_out/r2a_memcheck_autoreset.out
