#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
export MATE_B200_KERNEL=2
MATE_B200_LIB=/root/repo/scratch/variants/libmate_tl.so timeout 300 python scratch/timeline.py gpurun_out/r2h_timeline.npy 2>&1 | tail -120
run() { # name lib extra-args
  MATE_B200_LIB=/root/repo/scratch/variants/libmate_$2.so timeout 300 python bench.py --no-cpu --no-e2e ${@:3} > gpurun_out/r2h_$1.json 2>gpurun_out/r2h_$1.err
  python - <<PY
import json
try:
    d = json.load(open('gpurun_out/r2h_$1.json')); print('$1', round(d['ms_per_step'],5), round(d['roofline']['frac'],4), d['clocks']['sm_mhz'], d['episode_stats'])
except Exception as e: print('$1', 'FAILED', e, open('gpurun_out/r2h_$1.err').read()[-800:])
PY
}
run base_long base --steps 2000 --warmup 20
for i in 1 2 3 4 5; do run base_drv$i base --steps 20 --warmup 5; done
run p3_64k p3 --steps 1000 --warmup 20
run p3_32k p3 --steps 1000 --warmup 20 --envs 32768
run p3_16k p3 --steps 1000 --warmup 20 --envs 16384
run base_32k base --steps 1000 --warmup 20 --envs 32768
