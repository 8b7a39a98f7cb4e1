#!/bin/bash
cd /root/repo
timeout 1800 python -m pytest tests/test_cuda_parity.py tests/test_full_size_properties.py -m gpu -q -x 2>&1 | tail -4
for r in 1 2; do for v in lz0 lz1; do
MATE_B200_LIB=/root/repo/scratch/variants/libmate_$v.so timeout 300 python bench.py --steps 1000 --warmup 20 --no-cpu --no-e2e --no-configs | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$v 1000/20', d['ms_per_step'], d['roofline']['frac'])"
MATE_B200_LIB=/root/repo/scratch/variants/libmate_$v.so timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu --no-e2e --no-configs | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$v 20/5', d['ms_per_step'], d['roofline']['frac'])"
done; done
