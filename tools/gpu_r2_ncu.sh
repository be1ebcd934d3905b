#!/bin/bash
# Round-2 ncu evidence: (1) launch list of one steady-state step (short history), (2) --set full of every kernel family of
# one feval + the optimizers at full history.  Raw CSV exports come back in gpurun_out/<tag>/; tables are made by
# tools/ncu_table.py.   Usage (under gpurun): bash tools/gpu_r2_ncu.sh [tag]
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}"
TAG=${1:-r02n}
OUT=gpurun_out/$TAG
mkdir -p $OUT
BENCH="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-multires --no-extras"
T0=$(date +%s)
stamp() { echo "[$(( $(date +%s) - T0 )) s] $*"; }
export_rep() { ncu -i /tmp/$1.ncu-rep --page raw --csv > $OUT/$1_raw.csv 2>/dev/null; python tools/ncu_table.py $OUT/$1_raw.csv > $OUT/$1_table.txt 2>&1; cat $OUT/$1_table.txt; }

stamp "launch list (history prefill 10)"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file /tmp/launches.csv $BENCH --history-prefill 10 > $OUT/ncu_launches.log 2>&1
echo "exit $?"; python tools/launch_summary.py /tmp/launches.csv $OUT > $OUT/launch_summary.txt 2>&1; head -c 3000 $OUT/launch_summary.txt

stamp "ncu --set full: every conv_tc launch of one feval"
timeout 500 ncu --set full --clock-control none -k regex:conv_tc_kernel -s 72 -c 24 -o /tmp/conv_all -f $BENCH --history-prefill 0 > $OUT/ncu_conv.log 2>&1
echo "exit $?"; export_rep conv_all
python tools/ncu_table.py $OUT/conv_all_raw.csv --traffic 1024 > $OUT/conv_traffic.json; cat $OUT/conv_traffic.json

stamp "ncu --set full: conv_first / gram / pool / loss kernels of one feval"
timeout 400 ncu --set full --clock-control none -k regex:'gram_|conv_first|pool_|mse_|bwd_prep|loss_grad|style_' -s 60 -c 24 -o /tmp/misc_all -f $BENCH --history-prefill 0 > $OUT/ncu_misc.log 2>&1
echo "exit $?"; export_rep misc_all

stamp "ncu --set full with source: conv_first_fwd + 2 conv_tc (kept whole)"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'conv_first_fwd' -s 3 -c 1 -o $OUT/conv_first_src -f $BENCH --history-prefill 0 > $OUT/ncu_first_src.log 2>&1
echo "exit $?"; ls -la $OUT/*.ncu-rep

stamp "ncu --set full: L-BFGS at full history + Adam"
timeout 400 ncu --set full --clock-control none -k regex:'lbfgs_' -s 420 -c 8 -o /tmp/lbfgs_all -f $BENCH --history-prefill 100 > $OUT/ncu_lbfgs.log 2>&1
echo "exit $?"; export_rep lbfgs_all
timeout 200 ncu --set full --clock-control none -k regex:'adam_' -s 6 -c 3 -o /tmp/adam_all -f $BENCH --optimizer adam > $OUT/ncu_adam.log 2>&1
echo "exit $?"; export_rep adam_all
stamp done
du -sh $OUT
