#!/bin/bash
cd /root/repo
timeout 900 python -m pytest tests/test_cuda_parity.py -m gpu -x -q -k "step_host" 2>&1 | tail -2
for m in auto 0 auto 0 auto; do if [ $m = auto ]; then unset MATE_B200_HOST_COMPACT; else export MATE_B200_HOST_COMPACT=$m; fi; timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu --no-configs --e2e-steps 20 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('compact $m', d['e2e']['value'], d['roofline']['frac'])"; done
unset MATE_B200_HOST_COMPACT
MATE_B200_HOST_TRACE=1 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu --no-configs --e2e-steps 3 2>&1 | grep "step_host compact" | tail -1
nproc
