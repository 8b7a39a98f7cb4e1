#!/bin/bash
cd /root/repo
timeout 900 python scratch/debug_fullsize.py MATE-4v8-9.yaml 65536 2>&1 | tail -40
timeout 900 python scratch/debug_fullsize.py MATE-Navigation.yaml 65536 2>&1 | tail -40
timeout 1200 python -m pytest tests/test_fov_range.py -m gpu -q 2>&1 | tail -5
