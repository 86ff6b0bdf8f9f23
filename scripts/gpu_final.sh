#!/bin/bash
# Round-end validation on one GPU box: every GPU parity test, smoke, the bench line the driver will reproduce, the ncu launch list of
# one SERIAL step (SERL_PIPELINE=0: one step = one front end + one heads chain) with DRAM bytes per launch, and `--set full` captures
# of the dominant heads kernel and the sampler.  Everything lands in gpurun_out/ (copy what is judged into profiles/).
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
echo "== pytest gpu" ; timeout 1500 python -m pytest tests -m gpu -q --timeout=900 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1 ; echo "pytest rc=$?" ; tail -3 gpurun_out/pytest_gpu.log
echo "== smoke" ; timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1 ; echo "smoke rc=$?" ; tail -2 gpurun_out/smoke.log
echo "== bench (default flags)" ; timeout 900 python bench.py > gpurun_out/bench.log 2> gpurun_out/bench.err ; echo "bench rc=$?" ; tail -c 900 gpurun_out/bench.log ; tail -2 gpurun_out/bench.err
echo "== ncu launch list (+ DRAM bytes), serial step"
SERL_PIPELINE=0 SERL_BENCH_SKIP_SINGLE=1 SERL_BENCH_SKIP_CPU=1 timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r02.csv python bench.py --steps 2 --warmup 1 --sustain-s 0 > gpurun_out/ncu_bench_r02.log 2>&1 ; echo "ncu rc=$?"
echo "== ncu --set full (tgemm, sampler)"
SERL_PIPELINE=0 SERL_BENCH_SKIP_SINGLE=1 SERL_BENCH_SKIP_CPU=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:'tgemm_tf32_kernel|sample_frames_kernel|enc_finish_kernel' --launch-skip 13 -c 13 -o gpurun_out/r02_heads -f python bench.py --steps 2 --warmup 1 --sustain-s 0 > gpurun_out/ncu_full_r02.log 2>&1 ; echo "ncu full rc=$?" ; ls -la gpurun_out/*.ncu-rep | tail -3
