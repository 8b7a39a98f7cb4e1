#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
echo "== all gpu tests"
timeout 3000 python -m pytest tests -m gpu -q -s 2>&1 | grep -v "^$" | tail -15
