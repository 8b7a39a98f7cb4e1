#!/bin/bash
# round 2, GPU call 2: first run of the role-specialised step kernel (mate_step_kernel3): parity, then A/B timing
cd /root/repo
mkdir -p gpurun_out
echo "== smoke (kernel3)"
timeout 180 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
rc=${PIPESTATUS[0]}; echo "smoke rc=$rc"
if [ "$rc" = "124" ]; then echo "smoke timed out: stopping"; exit 1; fi
echo "== gpu tests (kernel3)"
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
echo "== parity tests with mate_step_kernel2 (refactored tile_body)"
MATE_B200_KERNEL=2 timeout 600 python -m pytest tests/test_cuda_parity.py tests/test_full_size_properties.py -m gpu -x -q 2>&1 | tail -5
echo "== A/B (ms/step, 1000 steps steady state)"
AB_ONE=1
for spec in k3a k3b k3c k3d k3a_tma k3b_tma k3a@MATE_B200_KERNEL=2 k3a; do
  n=${spec%%@*}; envs=""
  if [[ "$spec" == *@* ]]; then envs=$(echo "${spec#*@}" | tr ',' ' '); fi
  tag=$(echo "$spec" | tr '@=,' '___')
  env $envs MATE_B200_LIB=/root/repo/scratch/variants/libmate_$n.so timeout 300 python bench.py --no-cpu --no-e2e --steps 1000 --warmup 20 > gpurun_out/r2b_$tag.json 2>gpurun_out/r2b_$tag.err
  python - <<PY
import json
try:
    d = json.load(open('gpurun_out/r2b_$tag.json')); print('$spec', round(d['ms_per_step'],5), round(d['roofline']['frac'],4), d['clocks']['sm_mhz'], d['episode_stats'])
except Exception as e: print('$spec', 'FAILED', e, open('gpurun_out/r2b_$tag.err').read()[-800:])
PY
done
