#!/bin/bash
# First GPU visit of round 2 (≈ 8 minutes): verify the opt-in paths of DESIGN.md section 9, A/B them, and get the
# per-kernel picture at the small scales that dominate the multi-resolution job.
# Usage (under gpurun): bash tools/gpu_r2_first.sh [tag]
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}"
TAG=${1:-r02a}
OUT=gpurun_out/$TAG
mkdir -p $OUT
T0=$(date +%s)
stamp() { echo "[$(( $(date +%s) - T0 )) s] $*"; }
nvidia-smi --query-gpu=name,driver_version,memory.total,clocks.max.sm --format=csv > $OUT/gpu.txt 2>&1

stamp "1. variant goldens + opt-in paths (tests/test_zz_arch_gpu.py, MAUA_TEST_EXPERIMENTAL=1)"
MAUA_TEST_EXPERIMENTAL=1 timeout -k 5 240 python -m pytest tests/test_zz_arch_gpu.py -q -m gpu --timeout 120 -p no:cacheprovider -s \
  > $OUT/pytest_zz.log 2>&1
echo "exit $?"; grep -E "passed|failed|FAILED|Error|rel |PSNR|NACC" $OUT/pytest_zz.log | tail -70

bench() {  # name, env..., -- args
  local name=$1; shift
  local envs=()
  while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  env "${envs[@]}" timeout -k 5 120 python bench.py --no-cpu-baseline --no-multires "$@" --profile-out $OUT/profile_$name.json \
    > $OUT/bench_$name.json 2> $OUT/bench_$name.err
  echo "bench $name exit $?"
}
stamp "2. A/B at 1024^2 (L-BFGS, full history)"
bench base X=0 --
bench fusepool MAUA_FUSE_POOL=1 --
bench ffma2 MAUA_CONV1_FFMA2=1 --
bench nacc4 MAUA_GRAM_NACC=4 --
bench all MAUA_FUSE_POOL=1 MAUA_CONV1_FFMA2=1 MAUA_GRAM_NACC=4 --
stamp "3. small scales: where does an iteration go at 256^2 / 512^2"
bench 512 X=0 -- --size 512
bench 256 X=0 -- --size 256
bench 512_all MAUA_FUSE_POOL=1 MAUA_CONV1_FFMA2=1 -- --size 512
bench 256_all MAUA_FUSE_POOL=1 MAUA_CONV1_FFMA2=1 -- --size 256
bench 2048 X=0 -- --size 2048 --steps 20
bench adam1024 X=0 -- --optimizer adam
python - "$OUT" <<'PY'
import json, sys, glob
for f in sorted(glob.glob(sys.argv[1] + '/bench_*.json')):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('/')[-1].ljust(24), 'it/s %8.1f  ms %7.3f  e2e %8.1f  conv %6.1f TF/s (%.3f)  clk %s %s' % (
            d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['achieved'], d['roofline']['frac'],
            d['clocks']['sm_mhz'], d['clocks']['reasons']))
        print('   ', d['kernel_breakdown_ms'])
    except Exception as e:
        print(f, 'parse failed', e)
PY
stamp "3b. two / three images per GPU on separate streams (experimental, DESIGN.md section 6)"
for S in 2 3; do
  timeout -k 5 150 python bench.py --streams $S > $OUT/bench_streams$S.json 2> $OUT/bench_streams$S.err; echo "streams $S exit $?"; cat $OUT/bench_streams$S.json | cut -c1-400
done
stamp "4. launch list at 512^2 (short history so that the pass stays under 2 minutes)"
timeout -k 5 150 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file /tmp/launches_512.csv \
  python bench.py --size 512 --steps 2 --warmup 3 --no-cpu-baseline --no-multires --history-prefill 2 > $OUT/ncu_launches_512.log 2>&1
echo "exit $?"; python tools/launch_summary.py /tmp/launches_512.csv $OUT > $OUT/launch_summary_512.txt 2>&1; head -c 2500 $OUT/launch_summary_512.txt
stamp "5. whole suite + default bench (what the driver runs)"
timeout -k 5 200 python -m pytest tests -x -q -m gpu --timeout 200 -p no:cacheprovider > $OUT/pytest_gpu.log 2>&1
echo "exit $?"; tail -3 $OUT/pytest_gpu.log
timeout -k 5 200 python bench.py > $OUT/bench_default.json 2> $OUT/bench_default.err; echo "exit $?"; head -c 600 $OUT/bench_default.json
stamp done
