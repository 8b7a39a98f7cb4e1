#!/bin/bash
cd /root/repo
for k in 12 0; do for w in 5 20 60 200; do MATE_B200_L2_KEEP=$k python bench.py --steps 20 --warmup $w --no-cpu --no-e2e --no-configs | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print('keep $k warmup $w', round(d['ms_per_step'],5), round(d['roofline']['frac'],4))"; done; done
for k in 12 0; do MATE_B200_L2_KEEP=$k python bench.py --steps 200 --warmup 5 --no-cpu --no-e2e --no-configs | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print('keep $k steps 200 warmup 5', round(d['ms_per_step'],5), round(d['roofline']['frac'],4))"; done
