"""Turn the files profiles/tools/r2_profile.sh TAG left in gpurun_out/ into the committed evidence under profiles/:
raw metrics table, launch list, bench lines, stall table, and profiles/traffic.json tagged with the kernel-source hash.
usage: python profiles/tools/collect_profile.py TAG"""
import csv, hashlib, json, os, shutil, statistics, sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
tag = sys.argv[1]
src = lambda name: os.path.join(ROOT, 'gpurun_out', f'{tag}_{name}')
dst = lambda name: os.path.join(ROOT, 'profiles', f'{tag}_{name}')

rows = list(csv.reader(open(src('raw.csv'), newline='')))
header, units, values = rows[0], rows[1], rows[2]
table = {}
with open(dst('raw_metrics.txt'), 'w') as f:
    for name, unit, value in sorted(zip(header, units, values)):
        if '__' in name and value not in ('', 'n/a'):
            f.write(f'{name} {unit} {value}\n')
            table[name] = (unit, value)

def to_bytes(name):
    unit, value = table[name]
    scale = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}[unit]
    return float(value) * scale

h = hashlib.sha256()
for name in ('mate_step.cuh', 'mate_common.cuh'):
    h.update(open(os.path.join(ROOT, 'mate_b200', 'csrc', name), 'rb').read())
traffic_path = os.path.join(ROOT, 'profiles', 'traffic.json')
traffic = json.load(open(traffic_path))
traffic['_comment'] = traffic['_comment'].replace('r2u_raw_metrics.txt', f'{tag}_raw_metrics.txt')
traffic['MATE-4v8-9'] = {'bytes': int(to_bytes('dram__bytes_read.sum') + to_bytes('dram__bytes_write.sum')),
                         'kernel_sources_sha16': h.hexdigest()[:16],
                         'capture': f'profiles/{tag}_raw_metrics.txt (ncu --set full --clock-control none, profiles/tools/r2_profile.sh {tag})'}
json.dump(traffic, open(traffic_path, 'w'), indent=1)

launches = [r for r in csv.reader(open(src('launches.csv'), newline='')) if len(r) > 5 and r[0].isdigit()]
times = [(r[4], float(r[-1])) for r in launches if 'gpu__time_duration' in r[-3] or True]
ns = [t * (1000.0 if launches and launches[0][-2] in ('us', 'usecond') else 1.0) for _, t in times]
with open(dst('launches.txt'), 'w') as f:
    steps = ns[3:]
    f.write(f'ncu --metrics gpu__time_duration.sum --clock-control none -k regex:mate_step -c 70 (bench.py --steps 40 --warmup 3 --no-cpu --no-e2e --no-configs): {len(ns)} launches\n')
    f.write(f'first three (reset, prepare, first step after set_state: in-place work) {[int(x) for x in ns[:3]]} ns\n')
    f.write(f'step launches: n={len(steps)} median {statistics.median(steps) / 1e3:.1f} us min {min(steps) / 1e3:.1f} max {max(steps) / 1e3:.1f} (cold cache, serialised by the profiler)\n')
    for i, ((name, _), t) in enumerate(zip(times, ns)):
        f.write(f'{i} {name[:60]} {int(t)} ns\n')

shutil.copy(src('stalls_by_line.txt'), dst('stalls_by_line.txt'))
for a, b in (('bench_drv.json', 'bench_line.json'), ('bench_long.json', 'bench_line_2000_steps.json'), ('bench_ref.json', 'bench_line_reference_arm.json')):
    shutil.copy(src(a), dst(b))
print(open(dst('launches.txt')).read().split('\n')[2])
print(traffic['MATE-4v8-9'])
