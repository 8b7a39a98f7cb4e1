#!/bin/bash
cd /root/repo
timeout 900 python -m pytest tests/test_cuda_parity.py -m gpu -x -q -k "step_host" 2>&1 | tail -2
for m in default 0; do if [ $m = default ]; then unset MATE_B200_HOST_COMPACT; else export MATE_B200_HOST_COMPACT=$m; fi
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 --no-configs --e2e-steps 20 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('N=2 compact=$m', d['value'], d['roofline']['frac'], 'e2e', d['e2e']['value'])"; done
unset MATE_B200_HOST_COMPACT
CUDA_VISIBLE_DEVICES=0 python bench.py --steps 20 --warmup 5 --no-cpu --no-configs --e2e-steps 20 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('N=1 e2e', d['e2e']['value'])"
CUDA_VISIBLE_DEVICES=0 MATE_B200_BENCH_ROWS_KEPT=0 python bench.py --steps 20 --warmup 5 --no-cpu --no-configs --e2e-steps 20 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('N=1 full rewrite e2e', d['e2e']['value'])"
nproc
