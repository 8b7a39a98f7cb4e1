#!/usr/bin/env python
"""Top source lines per stall reason from an `ncu --page source --csv --print-source cuda,sass` dump.

usage: python profiles/ncu_stalls.py src.csv [reason=stall_long_sb] [top=15]
"""
import csv
import os
import sys


def main():
    path = sys.argv[1]
    reason = sys.argv[2] if len(sys.argv) > 2 else 'stall_long_sb'
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 15
    rows = list(csv.reader(open(path, encoding='utf-8', errors='replace')))
    hdr, fname, lines = None, '?', {}
    for r in rows:
        if r and r[0] == 'File Path':
            fname = os.path.basename(r[1]); continue
        if r and r[0] == 'Line No':
            hdr = r; continue
        if hdr is None or len(r) < len(hdr):
            continue
        if r[0] != '' and r[2] == '-':
            try:
                n = int(r[hdr.index(reason)])
            except ValueError:
                continue
            cur = lines.setdefault((fname, int(r[0])), [r[1].strip(), 0])
            cur[1] += n
    tot = sum(v[1] for v in lines.values()) or 1
    print(f'{reason}: {tot} samples')
    for k, v in sorted(lines.items(), key=lambda kv: -kv[1][1])[:top]:
        print(f'{k[0][:16]:16s}{k[1]:5d} {100 * v[1] / tot:5.1f}%  {v[0][:110]}')


if __name__ == '__main__':
    main()
