#!/bin/bash
cd /root/repo
export MATE_B200_LIB=/root/repo/scratch/variants/libmate_tlall.so
for spec in "MATE-8v8-9.yaml 32768" "MATE-Navigation.yaml 65536" "MATE-4v2-9.yaml 65536"; do set -- $spec
echo "=== $1 x $2"
TL_CONFIG=$1 TL_ENVS=$2 timeout 300 python scratch/timeline.py 2>&1 | tail -15
done
