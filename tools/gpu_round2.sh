#!/bin/bash
# One GPU-box visit (round 1, second half): every -m gpu test, bench lines at 1024 / 512 / 2048 / Adam, the reference arm,
# the ncu launch list and full-set captures.  Large .ncu-rep files stay in /tmp on the box; gpurun_out/ only receives
# CSV exports (gzipped), tables and one 3-launch report with source, so that it fits the 64 MiB return limit.
# Usage (under gpurun): bash tools/gpu_round2.sh [tag]
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}"
TAG=${1:-r01d}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,driver_version,memory.total,clocks.max.sm --format=csv > $OUT/gpu.txt 2>&1

echo "== pytest -m gpu (whole suite, one process, like the driver runs it) =="
MAUA_TEST_REPORT=$OUT/plan_report.txt timeout -k 10 700 python -m pytest tests -x -q -m gpu --timeout 300 -p no:cacheprovider -s > $OUT/pytest_gpu.log 2>&1
echo "exit $?"; grep -E "passed|failed|FAILED|Error|PSNR|cache" $OUT/pytest_gpu.log | tail -25

echo "== smoke =="
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "exit $?"; tail -2 $OUT/smoke.log

echo "== bench (default = 1024^2 L-BFGS) =="
timeout 900 python bench.py --profile-out $OUT/profile_1024.json > $OUT/bench_1024.json 2> $OUT/bench_1024.err
echo "exit $?"; tail -2 $OUT/bench_1024.err
for cfg in "512 lbfgs" "2048 lbfgs" "1024 adam" "256 adam"; do
  set -- $cfg
  timeout 600 python bench.py --size $1 --optimizer $2 --steps 30 --warmup 5 --no-cpu-baseline --no-multires > $OUT/bench_$1_$2.json 2> $OUT/bench_$1_$2.err
  echo "bench $1 $2 exit $?"
done
python - "$OUT" <<'PY'
import json, sys, glob
for f in sorted(glob.glob(sys.argv[1] + '/bench_*.json')):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('/')[-1], {k: (round(d[k], 3) if isinstance(d[k], float) else d[k]) for k in ['value', 'ms_per_step', 'gpu_launches']},
              'e2e', round(d['e2e']['value'], 2), 'roof', round(d['roofline']['achieved'], 1), round(d['roofline']['frac'], 3),
              'clk', d['clocks']['sm_mhz'], d['clocks']['reasons'])
        if 'multires_e2e' in d: print('  multires', d['multires_e2e'])
        if 'cpu_baseline' in d: print('  cpu', d['cpu_baseline'])
        print('  breakdown', d['kernel_breakdown_ms'])
        print('  hbm', d['hbm_kernels_gbs'])
    except Exception as e:
        print(f, 'parse failed', e)
PY
echo "== reference arm (bounded) =="
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/bench_ref.json 2> $OUT/bench_ref.err
echo "exit $?"; tail -c 400 $OUT/bench_ref.json

BENCH="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-multires"
echo "== ncu launch list (same command as the bench; L-BFGS at full history) =="
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -c 20000 --csv --log-file /tmp/launches_all.csv \
  $BENCH --history-prefill 100 > $OUT/ncu_launches.log 2>&1
echo "exit $?"; wc -l /tmp/launches_all.csv
python tools/launch_summary.py /tmp/launches_all.csv $OUT

export_rep() {  # name
  ncu -i /tmp/$1.ncu-rep --page raw --csv 2>/dev/null | gzip > $OUT/$1_raw.csv.gz
  ncu -i /tmp/$1.ncu-rep --page raw --csv > /tmp/$1_raw.csv 2>/dev/null
  python tools/ncu_table.py /tmp/$1_raw.csv > $OUT/$1_table.txt
  ls -la /tmp/$1.ncu-rep $OUT/$1_raw.csv.gz
}
echo "== ncu --set full: every conv_tc launch of one feval =="
timeout 900 ncu --set full --clock-control none -k regex:conv_tc_kernel -s 81 -c 27 -o /tmp/conv_all -f \
  $BENCH --history-prefill 0 > $OUT/ncu_conv.log 2>&1
echo "exit $?"; export_rep conv_all
python tools/ncu_table.py /tmp/conv_all_raw.csv --traffic 1024 > $OUT/conv_traffic.json; cat $OUT/conv_traffic.json
echo "== ncu --set full with source: 3 conv launches (kept whole) =="
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 90 -c 3 -o $OUT/conv_src -f \
  $BENCH --history-prefill 0 > $OUT/ncu_conv_src.log 2>&1
echo "exit $?"
echo "== ncu --set full: gram / edge / pool / optimizer / image kernels of one step at full history =="
timeout 900 ncu --set full --clock-control none -k regex:'gram_|lbfgs_|conv_first|pool_|adam|tv_|mse_|style_|resize|grid_sample|process' -s 4100 -c 45 -o /tmp/misc_all -f \
  $BENCH --history-prefill 100 > $OUT/ncu_misc.log 2>&1
echo "exit $?"; export_rep misc_all
rm -f $OUT/*.err.tmp; du -sh gpurun_out
