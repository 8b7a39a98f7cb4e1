#!/bin/bash
cd /root/repo
timeout 900 python -m pytest tests/test_cuda_parity.py -m gpu -x -q -k "step_host" 2>&1 | tail -2
for i in 1 2 3; do timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu --no-configs --e2e-steps 20 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('e2e', d['e2e']['value'], d['roofline']['frac'])"; done
MATE_B200_HOST_TRACE=1 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu --no-configs --e2e-steps 3 2>&1 | grep "step_host compact" | tail -2
