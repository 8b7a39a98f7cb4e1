#!/usr/bin/env python
"""Aggregate an `ncu --page source --csv --print-source cuda,sass` dump per CUDA source line
(per file), and per region of mate_step.cuh (regions are found by marker comments).

usage: ncu -i X.ncu-rep --page source --csv --print-source cuda,sass > src.csv
       python profiles/ncu_lines.py src.csv [top_n] [warp_tiles_per_launch]
"""
import csv
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
STEP_MARKS = [
    ('setup + env scalars', 'mate_step_kernel2(const Params p)'),
    ('camera simulate + derive', '--- _simulate (environment.py'),
    ('target simulate (fast)', '// Target.simulate (entities.py:645-668): fast path'),
    ('target simulate (exact requeue)', '// exact re-simulation of the targets near a disc'),
    ('aux / misc decl', 'uint32_t tdone_bits = 0;'),
    ('reset', '====== reset (environment.py'),
    ('sensing', '====== _update_view'),
    ('fov range+sector', '---- cameras: range + sector'),
    ('cc cache / cc mask', '// camera -> camera: the occlusion part is static'),
    ('mask store', 'if (view_active) {\n'),
    ('pair queue', '---- then the stochastic transmittance'),
    ('goals', '====== _assign_goals'),
    ('coverage', '// coverage statistics of the current view'),
    ('finish', '====== finish step'),
    ('write back', '--- write state back'),
    ('pack: loop + zero fill', '--- joint_observation (environment.py'),
    ('pack: entity scatter', 'for (int ep = 0; ep < EPASS; ++ep)'),
    ('pack: own rows', '// own rows: preserved data'),
    ('pack: bulk store', '// ---- staged rows -> HBM'),
]


def main():
    path = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    tiles = float(sys.argv[3]) if len(sys.argv) > 3 else 2048.0
    rows = list(csv.reader(open(path, encoding='utf-8', errors='replace')))
    hdr, fname, lines = None, '?', {}
    for r in rows:
        if r and r[0] == 'File Path':
            fname = os.path.basename(r[1])
            continue
        if r and r[0] == 'Line No':
            hdr = r
            continue
        if hdr is None or len(r) < len(hdr):
            continue
        if r[0] != '' and r[2] == '-':   # a source line summary row
            try:
                samples = int(r[hdr.index('# Samples')])
                inst = int(r[hdr.index('Instructions Executed')])
                thr = int(r[hdr.index('Thread Instructions Executed')])
            except ValueError:
                continue
            cur = lines.setdefault((fname, int(r[0])), [r[1].strip(), 0, 0, 0])
            cur[1] += samples
            cur[2] += inst
            cur[3] += thr
    tot_s = sum(v[1] for v in lines.values()) or 1
    tot_i = sum(v[2] for v in lines.values()) or 1
    print(f'total samples {tot_s}  total warp-instructions {tot_i}  ({tot_i / tiles:.0f} per warp tile)')
    # ---- per file
    files = {}
    for (f, _), v in lines.items():
        a = files.setdefault(f, [0, 0, 0])
        a[0] += v[1]; a[1] += v[2]; a[2] += v[3]
    for f, a in sorted(files.items(), key=lambda kv: -kv[1][1]):
        print(f'  {f:24s} inst {100 * a[1] / tot_i:5.1f}%  samples {100 * a[0] / tot_s:5.1f}%  lanes {a[2] / max(a[1], 1):4.1f}')
    # ---- regions of mate_step.cuh
    step_path = os.path.join(ROOT, 'mate_b200', 'csrc', 'mate_step.cuh')
    if os.path.exists(step_path) and any(f == 'mate_step.cuh' for f, _ in lines):
        src = open(step_path, encoding='utf-8').read().split('\n')
        starts = []
        for name, mark in STEP_MARKS:
            for i, l in enumerate(src):
                if mark.rstrip('\n') in l and (not starts or i + 1 > starts[-1][0]):
                    starts.append((i + 1, name))
                    break
        starts.sort()
        print('--- regions of mate_step.cuh')
        agg = {}
        for (f, ln), v in lines.items():
            if f != 'mate_step.cuh':
                continue
            name = 'helpers (above the kernel)'
            for s, n in starts:
                if ln >= s:
                    name = n
            a = agg.setdefault(name, [0, 0, 0])
            a[0] += v[1]; a[1] += v[2]; a[2] += v[3]
        order = ['helpers (above the kernel)'] + [n for _, n in starts]
        for n in order:
            if n in agg:
                a = agg[n]
                print(f'  {n:34s} inst {100 * a[1] / tot_i:5.1f}% ({a[1] / tiles:7.0f} per warp tile)  samples {100 * a[0] / tot_s:5.1f}%  lanes {a[2] / max(a[1], 1):4.1f}')
    print('--- by instructions executed')
    for k, v in sorted(lines.items(), key=lambda kv: -kv[1][2])[:top]:
        print(f'{k[0][:16]:16s}{k[1]:5d} inst {100 * v[2] / tot_i:5.1f}%  samples {100 * v[1] / tot_s:5.1f}%  lanes {v[3] / max(v[2], 1):4.1f}  {v[0][:100]}')
    print('--- by stall samples')
    for k, v in sorted(lines.items(), key=lambda kv: -kv[1][1])[:top]:
        print(f'{k[0][:16]:16s}{k[1]:5d} samples {100 * v[1] / tot_s:5.1f}%  inst {100 * v[2] / tot_i:5.1f}%  {v[0][:100]}')


if __name__ == '__main__':
    main()
