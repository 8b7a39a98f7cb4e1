#!/bin/bash
cd /root/repo
timeout 1800 python -m pytest tests/test_wrappers.py tests/test_env_api.py tests/test_aux_wrappers.py tests/test_agents.py -m gpu -q -x 2>&1 | tail -15
timeout 600 python scratch/time_wrappers.py 2>&1 | tail -6
timeout 300 python bench.py --steps 1000 --warmup 20 --no-cpu --no-e2e --no-configs | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('bare 1000/20', d['ms_per_step'], d['roofline']['frac'])"
