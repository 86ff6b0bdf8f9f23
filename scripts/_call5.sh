set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_fused_heads_gpu.py -m gpu -q -x -s --timeout=600 -p no:cacheprovider > gpurun_out/pytest_fused.log 2>&1; echo "fused rc=$?"; grep -E "^step|passed|failed|Error|error|assert" gpurun_out/pytest_fused.log | tail -30
SERL_BENCH_SKIP_SINGLE=1 SERL_BENCH_SKIP_CPU=1 timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_r02b.csv python bench.py --steps 2 --warmup 1 --sustain-s 0 > gpurun_out/ncu_bench_r02b.log 2>&1 ; echo "ncu rc=$?"
