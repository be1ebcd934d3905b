#!/bin/bash
# Last ~3 GPU-minutes of round 1: (1) launch list of two steady-state steps at FULL L-BFGS history (the first ~5500
# launches are skipped, not profiled, so the run stays short), (2) the re-bounded stored-Gram test, (3) ncu --set full of
# the L-BFGS kernels at full history.
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}"
OUT=gpurun_out/${1:-r01g}
mkdir -p $OUT
T0=$(date +%s)
stamp() { echo "[$(( $(date +%s) - T0 )) s] $*"; }
BENCH="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-multires --history-prefill 100"
stamp "1. launch list at full history"
timeout -k 5 80 ncu --metrics gpu__time_duration.sum --clock-control none -s 5514 -c 175 --csv --log-file /tmp/launches_full.csv \
  $BENCH > $OUT/ncu_launches_full.log 2>&1
echo "exit $?"; wc -l /tmp/launches_full.csv; cp /tmp/launches_full.csv $OUT/launches_full.csv
python tools/launch_summary.py /tmp/launches_full.csv $OUT > $OUT/launch_summary.txt 2>&1; head -c 2500 $OUT/launch_summary.txt
stamp "2. stored-Gram test"
timeout -k 5 60 python -m pytest tests/test_fullsize_gpu.py -q -m gpu -k stored_gram -p no:cacheprovider -s > $OUT/pytest_gram.log 2>&1
echo "exit $?"; grep -E "passed|failed|rel " $OUT/pytest_gram.log
stamp "3. ncu --set full: L-BFGS kernels at full history"
timeout -k 5 90 ncu --set full --clock-control none -k regex:lbfgs_ -s 408 -c 8 -o /tmp/lbfgs_full -f $BENCH > $OUT/ncu_lbfgs.log 2>&1
echo "exit $?"
ncu -i /tmp/lbfgs_full.ncu-rep --page raw --csv > /tmp/lbfgs_raw.csv 2>/dev/null
gzip -c /tmp/lbfgs_raw.csv > $OUT/lbfgs_full_raw.csv.gz
python tools/ncu_table.py /tmp/lbfgs_raw.csv > $OUT/lbfgs_full_table.txt; cat $OUT/lbfgs_full_table.txt
stamp done
