#!/bin/bash
# ncu evidence for profiles/: launch list of the bench command + full-set captures of the dominant kernels.
# The .ncu-rep files are large (about 1.6 MB per launch with --set full): they are written to /tmp on the box, their
# raw / details pages are exported as CSV into gpurun_out/ (what comes back), and only a 3-launch report with source
# is kept whole.   Usage (under gpurun): bash tools/gpu_profile.sh [tag]
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}"
TAG=${1:-r01}
OUT=gpurun_out/prof_${TAG}
mkdir -p $OUT
BENCH="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-multires"

echo "== launch list (same command as the bench, L-BFGS at full history) =="
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -c 20000 --csv --log-file /tmp/launches_all.csv \
  $BENCH --history-prefill 100 > $OUT/ncu_launches.log 2>&1
echo "exit $?"; wc -l /tmp/launches_all.csv
# keep the last 2 timed steps + e2e + profile passes: the tail of the list (one step = ~75 launches); the full list is big
python - "$TAG" <<'PY'
import csv, sys, collections, json
tag = sys.argv[1]
rows = []
with open('/tmp/launches_all.csv') as f:
    lines = [l for l in f if l.startswith('"')]
rd = csv.reader(lines)
hdr = next(rd)
ki, vi, ui = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
for r in rd:
    v = float(r[vi].replace(',', ''))
    u = r[ui]
    ns = v * {'ns': 1, 'us': 1e3, 'ms': 1e6, 'nsecond': 1, 'usecond': 1e3, 'msecond': 1e6, 'second': 1e9}.get(u, 1)
    rows.append((r[ki], ns))
open(f'gpurun_out/prof_{tag}/launches_tail.csv', 'w').write('kernel,ns\n' + '\n'.join(f'"{k}",{v:.0f}' for k, v in rows[-1200:]))
# one steady-state step = the launches between two consecutive lbfgs_update kernels near the end of the list
idx = [i for i, (k, _) in enumerate(rows) if 'lbfgs_update' in k]
out = {'total_launches_profiled': len(rows)}
if len(idx) >= 3:
    a, b = idx[-3] + 1, idx[-2] + 1
    step = rows[a:b]
    by = collections.OrderedDict()
    for k, v in step:
        short = k.split('(')[0]
        d = by.setdefault(short, [0, 0.0])
        d[0] += 1; d[1] += v
    tot = sum(v for _, v in step)
    out['one_step'] = {'launches': len(step), 'sum_us': tot / 1e3,
                       'by_kernel': {k: {'n': n, 'us': round(v / 1e3, 1), 'share': round(v / tot, 4)} for k, (n, v) in sorted(by.items(), key=lambda kv: -kv[1][1])}}
open(f'gpurun_out/prof_{tag}/launch_summary.json', 'w').write(json.dumps(out, indent=1))
print(json.dumps(out, indent=1)[:3000])
PY

export_rep() {  # name
  ncu -i /tmp/$1.ncu-rep --page raw --csv > $OUT/$1_raw.csv 2>/dev/null
  ncu -i /tmp/$1.ncu-rep --page details --csv > $OUT/$1_details.csv 2>/dev/null
  ls -la /tmp/$1.ncu-rep $OUT/$1_raw.csv
}
echo "== ncu --set full: every conv_tc launch of one feval =="
timeout 900 ncu --set full --clock-control none -k regex:conv_tc_kernel -s 81 -c 27 -o /tmp/conv_all -f \
  $BENCH --history-prefill 0 > $OUT/ncu_conv.log 2>&1
echo "exit $?"; export_rep conv_all
echo "== ncu --set full with source: 3 conv launches (kept whole) =="
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 85 -c 3 -o $OUT/conv_src -f \
  $BENCH --history-prefill 0 > $OUT/ncu_conv_src.log 2>&1
echo "exit $?"; ls -la $OUT/*.ncu-rep
echo "== ncu --set full: gram / edge / pool / optimizer kernels of one step =="
timeout 900 ncu --set full --clock-control none -k regex:'gram_|lbfgs_|conv_first|pool_|adam|tv_|mse_|style_' -s 200 -c 40 -o /tmp/misc_all -f \
  $BENCH --history-prefill 12 > $OUT/ncu_misc.log 2>&1
echo "exit $?"; export_rep misc_all
du -sh gpurun_out
