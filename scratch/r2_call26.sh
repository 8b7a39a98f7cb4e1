#!/bin/bash
cd /root/repo
timeout 1800 python -m pytest tests/test_wrappers.py -m gpu -q -x 2>&1 | tail -3
timeout 600 python scratch/time_wrappers2.py 2>&1 | tail -4
timeout 600 python scratch/time_wrappers.py 2>&1 | tail -4
