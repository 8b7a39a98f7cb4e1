#!/bin/bash
cd /root/repo
timeout 280 compute-sanitizer --tool memcheck --error-exitcode 3 python profiles/tools/sanitize_hostpath.py > gpurun_out/r2y_memcheck_hostpath.txt 2>&1; echo "memcheck rc=$?"; tail -4 gpurun_out/r2y_memcheck_hostpath.txt
