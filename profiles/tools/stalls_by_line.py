"""Aggregate an `ncu --page source --csv --print-source cuda,sass` export per source line and stall reason.
usage: python profiles/tools/stalls_by_line.py export.csv [top_n]"""
import csv, sys, collections

path = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 12
rows = list(csv.reader(open(path, newline='')))
cur_file, header = None, None
per = collections.defaultdict(lambda: collections.Counter())   # reason -> (file, line, text) -> samples
total = collections.Counter()
inst = collections.Counter()
for r in rows:
    if len(r) == 2 and r[0] == 'File Path':
        cur_file = r[1].split('/')[-1]
        continue
    if r and r[0] == 'Line No':
        header = r
        cols = {name: i for i, name in enumerate(header)}
        stall_cols = [(name, i) for i, name in enumerate(header) if name.startswith('stall_') and 'Not Issued' not in name]
        continue
    if header is None or len(r) < len(header) or not r[0].strip().isdigit():
        continue
    if r[cols['Address']].strip() not in ('', '-'):     # a SASS row below a source line: the source row already carries the sums
        continue
    key = (cur_file, int(r[0]), r[1].strip()[:110])
    for name, i in stall_cols:
        try:
            v = int(r[i] or 0)
        except ValueError:
            v = 0
        if v:
            per[name][key] += v
            total[name] += v
    try:
        inst[key] += int(r[cols['Instructions Executed']] or 0)
    except ValueError:
        pass
grand = sum(total.values())
print('samples by reason:', ', '.join(f'{k[6:]} {v} ({100 * v / grand:.1f}%)' for k, v in total.most_common()))
for name, n in total.most_common(7):
    print(f'\n{name}: {n} samples')
    for (f, line, text), v in per[name].most_common(top):
        print(f'  {f:18s}{line:5d} {100 * v / n:5.1f}%  {text}')
print('\nwarp instructions by line:')
ti = sum(inst.values())
for (f, line, text), v in inst.most_common(top):
    print(f'  {f:18s}{line:5d} {100 * v / ti:5.1f}%  {text}')
