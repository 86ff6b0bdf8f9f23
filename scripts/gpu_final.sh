#!/bin/bash
# Round-end validation on one GPU box: full parity suite, smoke, the bench line, an ncu launch list of one step with DRAM
# bytes per launch, and one `--set full` capture of the dominant kernels.  Everything lands in gpurun_out/.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
echo "== pytest gpu" ; timeout 900 python -m pytest tests -m gpu -q --timeout=300 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1 ; echo "pytest rc=$?" ; tail -4 gpurun_out/pytest_gpu.log
echo "== smoke" ; timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1 ; echo "smoke rc=$?" ; tail -2 gpurun_out/smoke.log
echo "== bench" ; timeout 600 python bench.py --steps 100 --warmup 5 > gpurun_out/bench.log 2> gpurun_out/bench.err ; echo "bench rc=$?" ; tail -c 600 gpurun_out/bench.log ; tail -2 gpurun_out/bench.err
echo "== ncu launch list (+ DRAM bytes)"
SERL_BENCH_SKIP_DUAL=1 SERL_BENCH_SKIP_CPU=1 timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 > gpurun_out/ncu_bench.log 2>&1 ; echo "ncu rc=$?"
echo "== ncu --set full (dominant kernels)"
SERL_BENCH_SKIP_DUAL=1 SERL_BENCH_SKIP_CPU=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:'stem_tc_kernel|conv3x3_tc_kernel|gemm_tf32x3_kernel|pool_finish' --launch-skip 20 -c 12 -o gpurun_out/r01_top_kernels -f python bench.py --steps 2 --warmup 1 > gpurun_out/ncu_full.log 2>&1 ; echo "ncu full rc=$?" ; ls -la gpurun_out/*.ncu-rep
