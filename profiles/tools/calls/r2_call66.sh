#!/bin/bash
cd /root/repo
timeout 1400 python -m pytest tests/test_cuda_parity.py tests/test_full_size_properties.py -m gpu -x -q 2>&1 | tail -2
for k in 0 12 0 12; do MATE_B200_L2_KEEP=$k python bench.py --steps 200 --warmup 10 --no-cpu --no-e2e | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print('keep $k', round(d['ms_per_step'],5), round(d['roofline']['frac'],4), [(c['workload'].split()[0], round(c['ms_per_step'],4), round(c['frac'],3)) for c in d.get('configs',[])])"; done
python bench.py --steps 20 --warmup 5 --no-cpu --no-configs | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print('driver-like', round(d['ms_per_step'],5), round(d['roofline']['frac'],4), d['e2e']['value'])"
