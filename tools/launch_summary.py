"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: keep the tail of the list and the per-kernel
shares of one steady-state step (the launches between two consecutive lbfgs_update / adam kernels).
    python tools/launch_summary.py launches.csv outdir"""
import collections
import csv
import json
import sys

src, out = sys.argv[1], sys.argv[2]
with open(src) as f:
    lines = [l for l in f if l.startswith('"')]
rd = csv.reader(lines)
hdr = next(rd)
ki, vi, ui = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
rows = []
for r in rd:
    v = float(r[vi].replace(',', ''))
    ns = v * {'ns': 1, 'us': 1e3, 'ms': 1e6, 'nsecond': 1, 'usecond': 1e3, 'msecond': 1e6, 'second': 1e9}.get(r[ui], 1)
    rows.append((r[ki], ns))
open(f'{out}/launches_tail.csv', 'w').write('kernel,ns\n' + '\n'.join(f'"{k}",{v:.0f}' for k, v in rows[-1200:]))
idx = [i for i, (k, _) in enumerate(rows) if 'lbfgs_update' in k or 'adam_kernel' in k]
res = {'total_launches_profiled': len(rows)}
if len(idx) >= 3:
    a, b = idx[-3] + 1, idx[-2] + 1
    step = rows[a:b]
    by = collections.OrderedDict()
    for k, v in step:
        short = k.split('(')[0].replace('void ', '').replace('maua::<unnamed>::', '')
        d = by.setdefault(short, [0, 0.0])
        d[0] += 1
        d[1] += v
    tot = sum(v for _, v in step)
    res['one_step'] = {'launches': len(step), 'sum_us': round(tot / 1e3, 1),
                       'by_kernel': {k: {'n': n, 'us': round(v / 1e3, 1), 'share': round(v / tot, 4)}
                                     for k, (n, v) in sorted(by.items(), key=lambda kv: -kv[1][1])}}
    conv = sum(v for k, (n, v) in by.items() if 'conv_tc_kernel' in k)
    res['one_step']['conv_tc_share'] = round(conv / tot, 4)
open(f'{out}/launch_summary.json', 'w').write(json.dumps(res, indent=1))
print(json.dumps(res, indent=1)[:2500])
