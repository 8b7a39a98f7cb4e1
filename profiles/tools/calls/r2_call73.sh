#!/bin/bash
cd /root/repo
for t in 2 3 5 7 10; do MATE_B200_HOST_COMPACT=1 MATE_B200_HOST_THREADS=$t timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu --no-configs --e2e-steps 20 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('threads $t e2e', d['e2e']['value'])"; done
