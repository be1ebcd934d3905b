#!/bin/bash
# One short GPU-box visit (round 1, last 13 GPU-minutes): steps ordered by what only this visit can produce, every step
# bounded by its own timeout and writing straight into gpurun_out/ so that a cut-off call still returns what finished.
#   1. the new full-size parity / property tests + the image-side tests
#   2. ncu launch list of the bench command (current kernels, L-BFGS at full history)
#   3. ncu --set full of every conv_tc launch of one feval -> DRAM traffic per launch (bench.py roofline.traffic)
#   4. the bench line (default flags, as the driver runs it)
#   5. the rest of the -m gpu suite
#   6. ncu --set full of the Gram / pool / optimizer kernels
# Usage (under gpurun): bash tools/gpu_final.sh [tag]
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}"
TAG=${1:-r01f}
OUT=gpurun_out/$TAG
mkdir -p $OUT
T0=$(date +%s)
stamp() { echo "[$(( $(date +%s) - T0 )) s] $*"; }
nvidia-smi --query-gpu=name,driver_version,memory.total,clocks.max.sm --format=csv > $OUT/gpu.txt 2>&1

stamp "1. full-size + image tests"
timeout -k 5 200 python -m pytest tests/test_fullsize_gpu.py tests/test_image_gpu.py -q -m gpu --timeout 150 -p no:cacheprovider -s \
  > $OUT/pytest_new.log 2>&1
echo "exit $?"; grep -E "passed|failed|FAILED|Error|rel |PSNR" $OUT/pytest_new.log | tail -60

BENCH="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-multires"
stamp "2. ncu launch list"
timeout -k 5 170 ncu --metrics gpu__time_duration.sum --clock-control none -c 20000 --csv --log-file /tmp/launches_all.csv \
  $BENCH --history-prefill 100 > $OUT/ncu_launches.log 2>&1
echo "exit $?"; wc -l /tmp/launches_all.csv
python tools/launch_summary.py /tmp/launches_all.csv $OUT > $OUT/launch_summary.txt 2>&1; head -c 1500 $OUT/launch_summary.txt

export_rep() {  # name
  ncu -i /tmp/$1.ncu-rep --page raw --csv > /tmp/$1_raw.csv 2>/dev/null
  gzip -c /tmp/$1_raw.csv > $OUT/$1_raw.csv.gz
  python tools/ncu_table.py /tmp/$1_raw.csv > $OUT/$1_table.txt
}
stamp "3. ncu --set full: conv_tc launches of one feval"
timeout -k 5 170 ncu --set full --clock-control none -k regex:conv_tc_kernel -s 81 -c 27 -o /tmp/conv_all -f \
  $BENCH --history-prefill 0 > $OUT/ncu_conv.log 2>&1
echo "exit $?"; export_rep conv_all; cat $OUT/conv_all_table.txt
python tools/ncu_table.py /tmp/conv_all_raw.csv --traffic 1024 > $OUT/conv_traffic.json && cp $OUT/conv_traffic.json profiles/conv_traffic.json
cat $OUT/conv_traffic.json

stamp "4. bench (default flags)"
timeout -k 5 240 python bench.py --profile-out $OUT/profile_1024.json > $OUT/bench_1024.json 2> $OUT/bench_1024.err
echo "exit $?"; tail -3 $OUT/bench_1024.err; head -c 3000 $OUT/bench_1024.json

stamp "5. rest of the -m gpu suite"
timeout -k 5 400 python -m pytest tests -q -m gpu --timeout 200 -p no:cacheprovider -s \
  --deselect tests/test_fullsize_gpu.py --deselect tests/test_image_gpu.py > $OUT/pytest_rest.log 2>&1
echo "exit $?"; grep -E "passed|failed|FAILED|Error" $OUT/pytest_rest.log | tail -20

stamp "6. ncu --set full: gram / pool / optimizer kernels at full history"
# window = one steady-state step at full history, located in the launch list of step 2 (same command, same launch order)
read SKIP COUNT < <(python - <<'PY'
import csv, re
try:
    lines = [l for l in open('/tmp/launches_all.csv') if l.startswith('"')]
    rd = csv.reader(lines); hdr = next(rd); ki = hdr.index('Kernel Name')
    names = [r[ki].split('(')[0] for r in rd]
    pat = re.compile(r'gram_|lbfgs_|conv_first|pool_|tv_|mse_|style_')
    upd = [i for i, k in enumerate(names) if 'lbfgs_update' in k]
    a, b = upd[-3] + 1, upd[-2] + 1
    skip = sum(1 for k in names[:a] if pat.search(k))
    cnt = sum(1 for k in names[a:b] if pat.search(k))
    print(skip, cnt + 1)
except Exception:
    print(3000, 45)
PY
)
echo "misc window: skip $SKIP count $COUNT"
timeout -k 5 170 ncu --set full --clock-control none -k regex:'gram_|lbfgs_|conv_first|pool_|tv_|mse_|style_' -s $SKIP -c $COUNT -o /tmp/misc_all -f \
  $BENCH --history-prefill 100 > $OUT/ncu_misc.log 2>&1
echo "exit $?"; export_rep misc_all; cat $OUT/misc_all_table.txt

stamp "7. reference arm + smoke"
timeout -k 5 200 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/bench_ref.json 2> $OUT/bench_ref.err
echo "exit $?"; head -c 600 $OUT/bench_ref.json
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $OUT/smoke.log 2>&1; tail -1 $OUT/smoke.log
stamp done; du -sh gpurun_out
