#!/bin/bash
# One GPU-box round trip: parity tests, smoke, short bench, ncu launch list.  Everything lands in gpurun_out/.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
echo "== pytest gpu" ; timeout 900 python -m pytest tests -m gpu -q -x --timeout=300 -p no:cacheprovider ${PYTEST_ARGS:-} > gpurun_out/pytest_gpu.log 2>&1 ; echo "pytest rc=$?" ; tail -25 gpurun_out/pytest_gpu.log
echo "== smoke" ; timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1 ; echo "smoke rc=$?" ; tail -5 gpurun_out/smoke.log
echo "== bench" ; timeout 600 python bench.py --steps ${BENCH_STEPS:-10} --warmup 3 ${BENCH_ARGS:-} > gpurun_out/bench.log 2> gpurun_out/bench.err ; echo "bench rc=$?" ; tail -3 gpurun_out/bench.log ; tail -5 gpurun_out/bench.err
if [ "${NCU:-1}" = "1" ]; then
  echo "== ncu launch list"
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 ${BENCH_ARGS:-} > gpurun_out/ncu_bench.log 2>&1 ; echo "ncu rc=$?"
fi
