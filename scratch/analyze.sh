#!/bin/bash
# usage: scratch/analyze.sh TAG [top]   -- read gpurun_out/TAG_full.ncu-rep here
TAG=$1
cd /root/repo/gpurun_out
ncu -i ${TAG}_full.ncu-rep --page raw --csv > ${TAG}_raw.csv 2>/dev/null
ncu -i ${TAG}_full.ncu-rep --page source --csv --print-source cuda,sass > ${TAG}_src.csv 2>/dev/null
python /root/repo/profiles/ncu_metrics.py ${TAG}_raw.csv
python /root/repo/profiles/ncu_lines.py ${TAG}_src.csv ${2:-30} 2048
