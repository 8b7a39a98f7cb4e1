#!/bin/bash
cd /root/repo
MATE_B200_LIB=/root/repo/scratch/variants/libmate_tl.so timeout 300 python scratch/timeline.py gpurun_out/r2q_timeline.npy 2>&1 | tail -15
for i in 1 2 3; do timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu --no-e2e --no-configs | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('20/5', d['ms_per_step'], d['roofline']['frac'])"; done
