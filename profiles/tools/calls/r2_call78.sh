#!/bin/bash
cd /root/repo
LOCAL_WORLD_SIZE=8 MATE_B200_HOST_THREADS=3 python bench.py --steps 20 --warmup 5 --no-cpu --no-configs --e2e-steps 10 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('sharers 8 threads 3:', d['e2e']['value'], d['e2e']['d2h_leg'])"
MATE_B200_HOST_THREADS=3 python bench.py --steps 20 --warmup 5 --no-cpu --no-configs --e2e-steps 10 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('sharers 1 threads 3:', d['e2e']['value'], d['e2e']['d2h_leg'])"
timeout 600 python -m pytest tests/test_cuda_parity.py -m gpu -x -q -k "step_host" 2>&1 | tail -1
python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu --no-configs | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('default:', d['ms_per_step'], d['roofline']['frac'], d['e2e']['value'], d['e2e']['d2h_leg'])"
