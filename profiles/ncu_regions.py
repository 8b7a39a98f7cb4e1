#!/usr/bin/env python
"""Aggregate an `ncu --page source --csv --print-source cuda,sass` dump by code region of
mate_kernels.cuh (regions are found by marker comments, so they follow the current source).

usage: python profiles/ncu_regions.py src.csv [warp_tiles_per_launch]
"""
import csv
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MARKS = [
    ('helpers', None), ('philox', 'Philox4x32-10, same draw'), ('cargo', 'struct Cargo'), ('obstruct', 'struct StepVec'),
    ('fov polyline (sight_range_at)', 'struct RaySample'), ('fov_reach', 'fov_reach_exact(double cx'),
    ('occlusion_fast', 'int occlusion_fast('), ('occlusion_exact', 'bool occlusion_exact('),
    ('camera_derive', 'void camera_derive('), ('env_reset', 'struct ResetCfg'), ('write_aux', 'struct AuxArgs'),
    ('assign_obstacles', 'void assign_obstacles('), ('kernel: setup', 'mate_step_kernel(const Params p)'),
    ('kernel: load', '--- load state'), ('kernel: simulate', '--- _simulate'),
    ('kernel: mask decl / aux', '// column masks of my entities'), ('kernel: reset', '====== reset'),
    ('kernel: view sensors', '====== _update_view'), ('kernel: view fov prefilter', '---- cameras: range + sector first'),
    ('kernel: query loop', '---- then the stochastic transmittance'), ('kernel: goals', '====== _assign_goals'),
    ('kernel: finish', '====== finish step'), ('kernel: write back', '--- write state back'),
    ('kernel: pack', '--- joint_observation (environment.py:908-983)'), ('kernel: copy out', '--- staged rows -> HBM'),
]


def main():
    rows = list(csv.reader(open(sys.argv[1], encoding='utf-8', errors='replace')))
    tiles = float(sys.argv[2]) if len(sys.argv) > 2 else 16384.0
    hdr, lines = None, {}
    for r in rows:
        if r and r[0] == 'Line No':
            hdr = r
            continue
        if hdr is None or len(r) < len(hdr) or r[0] == '' or r[2] != '-':
            continue
        try:
            k = int(r[0])
            vals = [int(r[hdr.index(n)]) for n in ('# Samples', 'Instructions Executed', 'Thread Instructions Executed')]
        except ValueError:
            continue
        cur = lines.setdefault(k, [0, 0, 0])
        for i in range(3):
            cur[i] += vals[i]
    src = open(os.path.join(ROOT, 'mate_b200', 'csrc', 'mate_kernels.cuh'), encoding='utf-8').read().split('\n')
    marks = []
    for name, needle in MARKS:
        if needle is None:
            marks.append((name, 1))
            continue
        for i, text in enumerate(src):
            if needle in text:
                marks.append((name, i + 1))
                break
    marks.sort(key=lambda x: x[1])
    tot_i = sum(v[1] for v in lines.values()) or 1
    tot_s = sum(v[0] for v in lines.values()) or 1
    for n, (name, start) in enumerate(marks):
        end = marks[n + 1][1] if n + 1 < len(marks) else 10 ** 9
        sel = [v for k, v in lines.items() if start <= k < end]
        inst, samp, thr = sum(v[1] for v in sel), sum(v[0] for v in sel), sum(v[2] for v in sel)
        print(f'{name:32s} inst {100 * inst / tot_i:5.1f}%  ({inst / tiles:6.0f} per warp tile)  samples {100 * samp / tot_s:5.1f}%  lanes {thr / max(inst, 1):4.1f}')
    print(f'total warp instructions per warp tile: {tot_i / tiles:.0f}')


if __name__ == '__main__':
    main()
