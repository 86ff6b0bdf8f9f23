set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_heads_fused_ops_gpu.py tests/test_fused_heads_gpu.py tests/test_pipeline_gpu.py tests/test_b256_fp16_gpu.py -m gpu -q --timeout=600 -p no:cacheprovider > gpurun_out/pytest_new.log 2>&1; echo "rc=$?"; grep -E "passed|failed|Error|error|assert|FAILED" gpurun_out/pytest_new.log | tail -12
