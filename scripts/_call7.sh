set -u
mkdir -p gpurun_out
TESTS="tests/test_fused_heads_gpu.py" TOOLS=memcheck SAN_TIMEOUT=900 PYTEST_ARGS="-k cams1" bash scripts/sanitize.sh
grep -B5 -A25 "Invalid\|out of bounds\|misaligned" gpurun_out/sanitize_memcheck.log | head -80
