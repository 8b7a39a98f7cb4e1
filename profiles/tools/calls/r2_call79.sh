#!/bin/bash
cd /root/repo
python bench.py --gpus 1 --steps 20 --warmup 5 --no-configs > gpurun_out/final2_bench_drv.json 2> gpurun_out/final2_bench_drv.err; tail -3 gpurun_out/final2_bench_drv.err
python -c "
import json; d=json.load(open('gpurun_out/final2_bench_drv.json')); print(d['ms_per_step'], d['roofline']['frac'], {k:v for k,v in d['e2e'].items() if k!='note'}, d['cpu_baseline']['value'])"
