#!/bin/bash
cd /root/repo
for m in 1 2; do MATE_B200_HOST_COMPACT=$m timeout 900 python -m pytest tests/test_cuda_parity.py -m gpu -x -q -k "step_host" 2>&1 | tail -2; done
for m in 0 1 2 0 1 2; do MATE_B200_HOST_COMPACT=$m timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu --no-configs --e2e-steps 20 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('compact $m', d['e2e']['value'])"; done
for m in 1 2; do MATE_B200_HOST_TRACE=1 MATE_B200_HOST_COMPACT=$m timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu --no-configs --e2e-steps 4 2>&1 | grep "step_host compact" | tail -2; done
for t in 6 10; do for m in 1 2; do MATE_B200_HOST_COMPACT=$m MATE_B200_HOST_THREADS=$t timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu --no-configs --e2e-steps 20 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('mode $m threads $t', d['e2e']['value'])"; done; done
