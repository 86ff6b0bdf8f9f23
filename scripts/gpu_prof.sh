#!/bin/bash
# Profiling round trip for the fused trunk kernels: launch list (time + DRAM bytes) of one step with the given switches, and one
# `ncu --set full` capture of the new kernels from a trunk-only driver.  Everything lands in gpurun_out/.
set -u
mkdir -p gpurun_out
FLAGS="${FLAGS:-SERL_STEM_V2=1 SERL_RES_CONV=1 SERL_RES_S2=1}"
echo "== trunk only, events ($FLAGS)"
env $FLAGS python scripts/prof_trunk.py fp16 256 5 2>&1 | tail -1
python scripts/prof_trunk.py fp16 256 5 2>&1 | tail -1
echo "== ncu launch list"
env $FLAGS SERL_BENCH_SKIP_SINGLE=1 SERL_BENCH_SKIP_CPU=1 timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 1100 --csv --log-file gpurun_out/launches_new.csv python bench.py --steps 2 --warmup 1 --sustain-s 0 > gpurun_out/ncu_bench_new.log 2>&1 ; echo "ncu rc=$?"
echo "== ncu --set full (new kernels)"
env $FLAGS timeout 900 ncu --set full --clock-control none --import-source on -k regex:'conv3x3_res_kernel|conv3x3s2_res_kernel|stem2_tc_kernel' --launch-skip 9 -c 9 -o gpurun_out/r02_new_kernels -f python scripts/prof_trunk.py fp16 256 2 > gpurun_out/ncu_full_new.log 2>&1 ; echo "ncu full rc=$?" ; ls -la gpurun_out/*.ncu-rep | tail -3
