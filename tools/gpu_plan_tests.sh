#!/bin/bash
# End-to-end plan parity tests on the GPU box (golden vectors + oracle), with a hard timeout.
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}"
mkdir -p gpurun_out
rm -f gpurun_out/plan_report.txt
MAUA_TEST_REPORT=gpurun_out/plan_report.txt timeout -k 10 900 python -m pytest tests/test_plan_gpu.py -q --timeout 300 -p no:cacheprovider "$@" > gpurun_out/plan_tests.log 2>&1
echo "== plan tests: exit $? =="; tail -n 40 gpurun_out/plan_tests.log; echo "== report =="; cat gpurun_out/plan_report.txt
