#!/bin/bash
cd /root/repo
timeout 1200 python -m pytest tests/test_fov_range.py -m gpu -q -x -k "4v2-9_random" 2>&1 | tail -30
timeout 1200 python -m pytest tests/test_full_size_properties.py -m gpu -q -x -k "oracle_at_full_size and 4v8-9" 2>&1 | tail -30
timeout 1200 python -m pytest tests/test_full_size_properties.py -m gpu -q -x -k "oracle_at_full_size and Navigation" 2>&1 | tail -30
