#!/bin/bash
# usage: scratch/final_cycle.sh TAG  -- full bench line, reference arm, other shapes, launch list
TAG=$1
python bench.py > gpurun_out/${TAG}_bench_full.json 2> gpurun_out/${TAG}_bench_full.err
python bench.py --impl reference --steps 5 --warmup 2 > gpurun_out/${TAG}_bench_ref.json 2> gpurun_out/${TAG}_bench_ref.err
for cfg in MATE-4v8-0 MATE-8v8-9 MATE-Navigation MATE-4v2-9; do
  python bench.py --config $cfg.yaml --no-cpu --no-e2e --steps 500 > gpurun_out/${TAG}_bench_$cfg.json 2> gpurun_out/${TAG}_bench_$cfg.err
done
python bench.py --config MATE-8v8-9.yaml --envs 32768 --no-cpu --no-e2e --steps 500 > gpurun_out/${TAG}_bench_MATE-8v8-9_32k.json 2>/dev/null
python bench.py --no-stagger --no-cpu --no-e2e --steps 1000 > gpurun_out/${TAG}_bench_nostagger.json 2>/dev/null
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:mate_step -c 70 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 40 --warmup 3 --no-cpu --no-e2e > gpurun_out/${TAG}_launches.log 2>&1
python - <<PY
import json, glob
for f in sorted(glob.glob('gpurun_out/${TAG}_bench_*.json')):
    try:
        d = json.load(open(f)); print(f.split('/')[-1], d.get('impl','mine'), d['config'].get('workload'), round(d['ms_per_step'],4), d.get('roofline',{}).get('frac'), d.get('e2e',{}).get('value'), d.get('cpu_baseline',{}).get('value'))
    except Exception as e: print(f, 'FAILED', e)
PY
