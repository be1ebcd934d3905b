#!/bin/bash
# Round-2 (second session) GPU visit: full parity suite, the bench line with the new legs, the multi-res job trace, and ncu
# evidence for the kernels added this session (conv_gen / pool3 / window SYRK) plus a refreshed launch list.
# Usage (under gpurun): bash tools/gpu_r3_all.sh [tag] [phases]     phases: any of t (tests) b (bench) m (multires trace) n (ncu)
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}"
TAG=${1:-r03f}; PH=${2:-tbmn}
OUT=gpurun_out/$TAG; mkdir -p $OUT
T0=$(date +%s); stamp() { echo "[$(( $(date +%s) - T0 )) s] $*"; }
nvidia-smi --query-gpu=name,driver_version,clocks.max.sm --format=csv > $OUT/gpu.txt 2>&1

if [[ $PH == *t* ]]; then
  stamp "pytest -m gpu"
  timeout 1500 python -m pytest tests -q -m gpu -s -p no:cacheprovider > $OUT/tests.log 2>&1; echo "pytest exit $?"
  tail -5 $OUT/tests.log
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
        print(k, ":", json.dumps(v)[:700])
except Exception as e:
    print("bench parse failed:", e)
PY
  tail -3 $OUT/bench.err
  stamp "reference arm (short)"
  timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/bench_reference.json 2> $OUT/bench_reference.err; echo "ref exit $?"; cut -c1-600 $OUT/bench_reference.json
fi
if [[ $PH == *m* ]]; then
  stamp "multi-res job phase trace"
  MAUA_TRACE=1 timeout 300 python tools/trace_multires.py > $OUT/multires_trace.txt 2>&1; echo "trace exit $?"; tail -40 $OUT/multires_trace.txt
fi
if [[ $PH == *n* ]]; then
  BENCH="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-multires --no-extras"
  export_rep() { ncu -i /tmp/$1.ncu-rep --page raw --csv > $OUT/$1_raw.csv 2>/dev/null; python tools/ncu_table.py $OUT/$1_raw.csv > $OUT/$1_table.txt 2>&1; cat $OUT/$1_table.txt; }
  stamp "launch list (history prefill 10)"
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file /tmp/launches.csv $BENCH --history-prefill 10 > $OUT/ncu_launches.log 2>&1
  echo "exit $?"; python tools/launch_summary.py /tmp/launches.csv $OUT > $OUT/launch_summary.txt 2>&1; head -c 2500 $OUT/launch_summary.txt
  stamp "ncu --set full: NIN feval at 2048^2 (conv_gen, pool3, pointwise conv_tc)"
  timeout 400 ncu --set full --clock-control none -k regex:'conv_gen|pool3|conv_tc|gram_tc' -c 60 -o /tmp/nin_all -f python tools/run_feval.py nin 2048 > $OUT/ncu_nin.log 2>&1
  echo "exit $?"; export_rep nin_all
  stamp "ncu --set full: pruned VGG-16 feval at 2048^2"
  timeout 400 ncu --set full --clock-control none -k regex:'conv_tc|gram_tc|conv_first' -c 40 -o /tmp/prune_all -f python tools/run_feval.py prune 2048 > $OUT/ncu_prune.log 2>&1
  echo "exit $?"; export_rep prune_all
  stamp "ncu --set full: img_vid window (4 frames, 512^2): SYRK over B*C channels + folded dgrad"
  timeout 400 ncu --set full --clock-control none -k regex:'gram_tc|gram_finalize' -c 12 -o /tmp/window_all -f python tools/run_feval.py window 512 > $OUT/ncu_window.log 2>&1
  echo "exit $?"; export_rep window_all
fi
stamp done; du -sh $OUT
