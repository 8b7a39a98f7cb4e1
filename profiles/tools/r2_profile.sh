#!/bin/bash
# usage: profiles/tools/r2_profile.sh TAG -- bench line (driver's command), reference arm, ncu full capture of the step kernel, launch list
TAG=$1
cd /root/repo; mkdir -p gpurun_out
python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/${TAG}_bench_drv.json 2> gpurun_out/${TAG}_bench_drv.err
python bench.py > gpurun_out/${TAG}_bench_long.json 2> gpurun_out/${TAG}_bench_long.err
python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/${TAG}_bench_ref.json 2> gpurun_out/${TAG}_bench_ref.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:mate_step_kernel2 -s 10 -c 1 -f -o /tmp/${TAG}_full python bench.py --steps 20 --warmup 3 --no-cpu --no-e2e --no-configs > gpurun_out/${TAG}_ncu.log 2>&1
ncu -i /tmp/${TAG}_full.ncu-rep --page raw --csv > gpurun_out/${TAG}_raw.csv 2>/dev/null
ncu -i /tmp/${TAG}_full.ncu-rep --page source --csv --print-source cuda,sass > /tmp/${TAG}_src.csv 2>/dev/null
python profiles/tools/stalls_by_line.py /tmp/${TAG}_src.csv 20 > gpurun_out/${TAG}_stalls_by_line.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:mate_step -c 70 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 40 --warmup 3 --no-cpu --no-e2e --no-configs > gpurun_out/${TAG}_launches.log 2>&1
python - <<PY
import json
for n in ('drv','long','ref'):
    try:
        d = json.load(open('gpurun_out/${TAG}_bench_%s.json' % n)); print(n, d.get('impl','mine'), round(d['ms_per_step'],5), d.get('roofline',{}).get('frac'), d.get('e2e',{}).get('value'), d.get('cpu_baseline',{}).get('value'), [(c['workload'], round(c['frac'],3)) for c in d.get('configs',[])])
    except Exception as e: print(n, 'FAILED', e)
PY
tail -2 gpurun_out/${TAG}_ncu.log
