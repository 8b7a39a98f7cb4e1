#!/usr/bin/env python
"""Aggregate an `ncu --page source --csv --print-source cuda,sass` dump per CUDA source line.

usage: ncu -i X.ncu-rep --page source --csv --print-source cuda,sass > src.csv
       python profiles/ncu_lines.py src.csv [top_n]
"""
import csv
import sys


def main():
    path = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    rows = list(csv.reader(open(path, encoding='utf-8', errors='replace')))
    hdr = None
    lines = {}
    for r in rows:
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
            key = int(r[0])
            cur = lines.setdefault(key, [r[1].strip(), 0, 0, 0])
            cur[1] += samples
            cur[2] += inst
            cur[3] += thr
    tot_s = sum(v[1] for v in lines.values()) or 1
    tot_i = sum(v[2] for v in lines.values()) or 1
    print(f'total samples {tot_s}  total warp-instructions {tot_i}')
    print('--- by instructions executed')
    for k, v in sorted(lines.items(), key=lambda kv: -kv[1][2])[:top]:
        print(f'{k:5d} inst {100 * v[2] / tot_i:5.1f}%  samples {100 * v[1] / tot_s:5.1f}%  lanes {v[3] / max(v[2], 1):4.1f}  {v[0][:110]}')
    print('--- by stall samples')
    for k, v in sorted(lines.items(), key=lambda kv: -kv[1][1])[:top]:
        print(f'{k:5d} samples {100 * v[1] / tot_s:5.1f}%  inst {100 * v[2] / tot_i:5.1f}%  {v[0][:110]}')


if __name__ == '__main__':
    main()
