#!/bin/bash
cd /root/repo
for m in 2 0; do MATE_B200_HOST_COMPACT=$m python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 2953$m bench.py --gpus 8 --steps 20 --warmup 5 --no-configs --e2e-steps 10 2>/dev/null | grep "^{" | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('N=8 compact=$m', d['value'], 'e2e', d['e2e']['value'], d['e2e']['d2h_leg'])"; done
