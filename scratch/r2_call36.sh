#!/bin/bash
cd /root/repo
timeout 1800 python -m pytest tests/test_aux_wrappers.py -m gpu -q -x 2>&1 | tail -5
timeout 600 python scratch/time_extras.py 2>&1 | tail -8
