#!/bin/bash
cd /root/repo
for v in oa2 oa4 oa6; do for spec in "MATE-4v2-9.yaml 65536" "MATE-Navigation.yaml 65536" "MATE-8v8-9.yaml 32768"; do set -- $spec
MATE_B200_LIB=/root/repo/scratch/variants/libmate_$v.so timeout 300 python bench.py --config $1 --envs $2 --steps 500 --warmup 20 --no-cpu --no-e2e --no-configs | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$v', '$1', round(d['ms_per_step'],5), round(d['roofline']['frac'],4))"
done; done
