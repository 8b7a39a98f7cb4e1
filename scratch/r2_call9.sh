#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
export MATE_B200_KERNEL=2
MATE_B200_LIB=/root/repo/scratch/variants/libmate_tl.so timeout 300 python scratch/timeline.py gpurun_out/r2h_timeline.npy 2>&1 | tail -45
