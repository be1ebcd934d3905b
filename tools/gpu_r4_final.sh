#!/bin/bash
# Round-2 (third session) GPU visit: full parity suite, the bench line with every leg + the reference arm, the multi-res job trace,
# warm-cache launch lists at 256^2 / 512^2 / 1024^2 and ncu --set full of the conv family (merged pipeline stages) and of the
# small-vector L-BFGS kernels.   bash tools/gpu_r4_final.sh [tag] [phases: t b m l n]
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}"
TAG=${1:-r04z}; PH=${2:-tbmln}
OUT=gpurun_out/$TAG; mkdir -p $OUT
T0=$(date +%s); stamp() { echo "[$(( $(date +%s) - T0 )) s] $*"; }
nvidia-smi --query-gpu=name,driver_version,clocks.max.sm --format=csv > $OUT/gpu.txt 2>&1
if [[ $PH == *t* ]]; then
  stamp "pytest -m gpu"
  timeout 1500 python -m pytest tests -q -m gpu -s -p no:cacheprovider > $OUT/tests.log 2>&1; echo "pytest exit $?"
  tail -3 $OUT/tests.log
  grep -E "^(nin|vgg16p|img_vid|exact)|PSNR| rel " $OUT/tests.log > $OUT/parity.txt
fi
if [[ $PH == *b* ]]; then
  stamp "bench (default line with all legs)"
  timeout 900 python bench.py --profile-out $OUT/prof_1024.json > $OUT/bench.json 2> $OUT/bench.err; echo "bench exit $?"
  python - $OUT/bench.json <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    for k, v in d.items():
        print(k, ":", json.dumps(v)[:600])
except Exception as e:
    print("bench parse failed:", e)
PY
  tail -3 $OUT/bench.err
  stamp "reference arm (short)"
  timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/bench_reference.json 2> $OUT/bench_reference.err; echo "ref exit $?"; cut -c1-400 $OUT/bench_reference.json
fi
if [[ $PH == *m* ]]; then
  stamp "multi-res job phase trace"
  MAUA_TRACE=1 timeout 300 python tools/trace_multires.py > $OUT/multires_trace.txt 2>&1; echo "trace exit $?"; tail -25 $OUT/multires_trace.txt
fi
if [[ $PH == *l* ]]; then
  stamp "launch lists (warm caches, history prefill 8)"
  bash tools/gpu_r4_launches.sh $TAG "256 512 1024" > $OUT/launch_lists.txt 2>&1; grep -E "sum_us|conv_tc_share" $OUT/l*/launch_summary.json
fi
if [[ $PH == *n* ]]; then
  BENCH="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-multires --no-extras --history-prefill 8"
  export_rep() { ncu -i /tmp/$1.ncu-rep --page raw --csv > $OUT/$1_raw.csv 2>/dev/null; python tools/ncu_table.py $OUT/$1_raw.csv > $OUT/$1_table.txt 2>&1; cat $OUT/$1_table.txt; gzip -f $OUT/$1_raw.csv; }
  stamp "ncu --set full: conv_tc of one iteration at 1024^2"
  timeout 500 ncu --set full --clock-control none -k regex:'conv_tc_kernel|conv_first' -s 120 -c 27 -o /tmp/conv_all -f $BENCH > $OUT/ncu_conv.log 2>&1
  echo "exit $?"; export_rep conv_all
  stamp "ncu --set full: conv_tc of one iteration at 256^2"
  timeout 400 ncu --set full --clock-control none -k regex:'conv_tc_kernel' -s 120 -c 24 -o /tmp/conv_256 -f $BENCH --size 256 > $OUT/ncu_conv256.log 2>&1
  echo "exit $?"; export_rep conv_256
  stamp "ncu --set full: L-BFGS kernels at full history, 256^2 (small-vector kernels) and 1024^2"
  timeout 300 ncu --set full --clock-control none -k regex:lbfgs_ -s 440 -c 4 -o /tmp/lbfgs_256 -f python bench.py --size 256 --steps 8 --warmup 3 --no-cpu-baseline --no-multires --no-extras > $OUT/ncu_lbfgs256.log 2>&1
  echo "exit $?"; export_rep lbfgs_256
  timeout 400 ncu --set full --clock-control none -k regex:lbfgs_ -s 440 -c 4 -o /tmp/lbfgs_1024 -f python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-multires --no-extras > $OUT/ncu_lbfgs1024.log 2>&1
  echo "exit $?"; export_rep lbfgs_1024
fi
stamp done; du -sh $OUT
