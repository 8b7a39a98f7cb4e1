#!/bin/bash
cd /root/repo
for v in fx0 fx1 fx2 fx3; do echo "== $v"; MATE_B200_LIB=/root/repo/scratch/variants/libmate_$v.so timeout 600 python scratch/time_wrappers2.py 2>&1 | tail -4; done
