#!/bin/bash
cd /root/repo
for b in 2 3 4 6 100; do MATE_B200_HOST_DENSE_EVERY=$b timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu --no-configs --e2e-steps 20 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('dense every $b', d['e2e']['value'])"; done
for b in 3 4; do MATE_B200_HOST_TRACE=1 MATE_B200_HOST_DENSE_EVERY=$b timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu --no-configs --e2e-steps 4 2>&1 | grep "step_host compact" | tail -2; done
for t in 3 7; do MATE_B200_HOST_DENSE_EVERY=2 MATE_B200_HOST_THREADS=$t timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu --no-configs --e2e-steps 20 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('every 2 threads $t', d['e2e']['value'])"; done
