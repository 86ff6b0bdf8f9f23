#!/bin/bash
# compute-sanitizer over the small-shape GPU tests: memcheck on every kernel family of the hot path, racecheck on
# the shared-memory pipelines.  Run on a GPU box: `gpurun --timeout 1800 -- 'bash scripts/sanitize.sh'`; logs in gpurun_out/.
set -u
mkdir -p gpurun_out
CS=${CS:-/usr/local/cuda/bin/compute-sanitizer}
TESTS=${TESTS:-"tests/test_tgemm_gpu.py tests/test_fused_heads_gpu.py tests/test_ops_gpu.py tests/test_replay_device.py"}
for tool in ${TOOLS:-memcheck}; do
  echo "== compute-sanitizer --tool $tool"
  timeout ${SAN_TIMEOUT:-1500} $CS --tool $tool --error-exitcode 77 --launch-timeout 120 python -m pytest $TESTS -m gpu -q -x --timeout=1200 -p no:cacheprovider ${PYTEST_ARGS:-} > gpurun_out/sanitize_$tool.log 2>&1
  echo "$tool rc=$?"; grep -E "ERROR SUMMARY|passed|failed|Invalid|Error:|at .*\.cu" gpurun_out/sanitize_$tool.log | head -30
done
