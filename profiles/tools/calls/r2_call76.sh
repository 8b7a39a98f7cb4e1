#!/bin/bash
cd /root/repo
timeout 1400 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/final_bench_drv.json 2> gpurun_out/final_bench_drv.err
python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/final_bench_ref.json 2> gpurun_out/final_bench_ref.err
python - <<PY
import json
d=json.load(open('gpurun_out/final_bench_drv.json')); r=json.load(open('gpurun_out/final_bench_ref.json'))
print(d['ms_per_step'], d['roofline']['frac'], d['roofline']['traffic_matches_this_build'], d['e2e'], d['cpu_baseline']['value'], r['value'], d['clocks'])
PY
