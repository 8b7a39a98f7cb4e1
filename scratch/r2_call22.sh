#!/bin/bash
cd /root/repo
timeout 1800 python -m pytest tests/test_wrappers.py tests/test_env_api.py -m gpu -q -x 2>&1 | tail -15
timeout 600 python scratch/time_wrappers.py 2>&1 | tail -6
