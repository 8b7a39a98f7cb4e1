#!/bin/bash
cd /root/repo
MATE_B200_HOST_BACKLOG=50 timeout 900 python -m pytest tests/test_cuda_parity.py -m gpu -x -q -k "step_host" 2>&1 | tail -2
for b in 60 100 150 200 300; do MATE_B200_HOST_BACKLOG=$b timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu --no-configs --e2e-steps 20 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('backlog $b', d['e2e']['value'])"; done
for b in 60 100 150; do MATE_B200_HOST_TRACE=1 MATE_B200_HOST_BACKLOG=$b timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu --no-configs --e2e-steps 4 2>&1 | grep "step_host compact" | tail -2; done
for t in 1 3 7; do MATE_B200_HOST_THREADS=$t timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu --no-configs --e2e-steps 20 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('threads $t', d['e2e']['value'])"; done
