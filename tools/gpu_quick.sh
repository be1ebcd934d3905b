#!/bin/bash
# Quick GPU check: all -m gpu tests (grouped, with timeouts) + one bench line.  Usage: bash tools/gpu_quick.sh [bench args]
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}"
mkdir -p gpurun_out
bash tools/gpu_kernel_tests.sh 2>&1 | grep -E "^==|passed|failed|FAILED|Error"
bash tools/gpu_plan_tests.sh 2>&1 | grep -E "^==|passed|failed|FAILED|Error|gradient rel vs"
timeout 600 python -m pytest tests/test_parallel.py -q -m gpu --timeout 300 -p no:cacheprovider -s 2>&1 | grep -E "passed|failed|FAILED|Error|staged" | tail -20
timeout 900 python bench.py --steps 30 --warmup 5 --profile-out gpurun_out/profile_1024.json "$@" > gpurun_out/bench_1024.json 2> gpurun_out/bench_1024.err
echo "== bench exit $? =="; tail -3 gpurun_out/bench_1024.err
python - <<'PY'
import json
try:
    d = json.loads(open('gpurun_out/bench_1024.json').read().strip().splitlines()[-1])
    print({k: d[k] for k in ['value', 'ms_per_step', 'gpu_launches', 'clocks']})
    print('e2e', d['e2e'])
    print('roofline', {k: d['roofline'][k] for k in ['achieved', 'peak', 'frac', 'share_of_feval']})
    print('breakdown', d['kernel_breakdown_ms'])
    print('cpu', d.get('cpu_baseline'))
except Exception as e:
    print('bench parse failed', e)
PY
