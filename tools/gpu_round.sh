#!/bin/bash
# One GPU-box visit: parity tests, bench, ncu launch list and one full ncu capture of the dominant kernels.
# Usage (under gpurun): bash tools/gpu_round.sh [tag]
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}"
TAG=${1:-r01}
mkdir -p gpurun_out
bash tools/gpu_all.sh
echo "== reference arm (bounded) =="
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
tail -c 600 gpurun_out/bench_ref.json
echo "== ncu launch list =="
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 400 --csv --log-file gpurun_out/launches_${TAG}.csv \
  python bench.py --steps 2 --warmup 3 --history-prefill 2 --no-cpu-baseline --no-multires > gpurun_out/ncu_bench.log 2>&1
echo "exit $?"; wc -l gpurun_out/launches_${TAG}.csv
echo "== ncu full: conv kernels inside the iteration =="
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 60 -c 25 -o gpurun_out/prof_conv_${TAG} -f \
  python bench.py --steps 1 --warmup 3 --history-prefill 0 --no-cpu-baseline --no-multires > gpurun_out/ncu_conv.log 2>&1
echo "exit $?"; ls -la gpurun_out/*.ncu-rep
echo "== ncu full: gram + memory-bound kernels =="
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'gram_tc|lbfgs_|conv_first|pool_|adam' -s 40 -c 24 -o gpurun_out/prof_misc_${TAG} -f \
  python bench.py --steps 1 --warmup 3 --history-prefill 4 --no-cpu-baseline --no-multires > gpurun_out/ncu_misc.log 2>&1
echo "exit $?"; ls -la gpurun_out/*.ncu-rep
