#!/bin/bash
# One comprehensive GPU-box round trip for kernel work: (1) parity tests, default path + every new switch on, (2) A/B bench lines,
# (3) ncu launch list (time + DRAM bytes) of one step with the switches in $FLAGS, (4) `ncu --set full` of the fused trunk kernels
# from a trunk-only driver.  Everything lands in gpurun_out/.
set -u
mkdir -p gpurun_out
FLAGS="${FLAGS:-SERL_STEM_V2=1 SERL_RES_CONV=1 SERL_RES_S2=1 SERL_SAMPLER_PERSISTENT=1}"
NCU=0 AB=1 BENCH=0 bash scripts/gpu_check.sh
echo "== trunk only, CUDA events: new switches, then default"
env $FLAGS python scripts/prof_trunk.py fp16 256 5 2>&1 | tail -1
python scripts/prof_trunk.py fp16 256 5 2>&1 | tail -1
echo "== ncu launch list ($FLAGS)"
env $FLAGS SERL_BENCH_SKIP_SINGLE=1 SERL_BENCH_SKIP_CPU=1 timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 1100 --csv --log-file gpurun_out/launches_new.csv python bench.py --steps 2 --warmup 1 --sustain-s 0 > gpurun_out/ncu_bench_new.log 2>&1 ; echo "ncu rc=$?"
echo "== ncu --set full (new kernels)"
env $FLAGS timeout 900 ncu --set full --clock-control none --import-source on -k regex:'conv3x3_res_kernel|conv3x3s2_res_kernel|stem2_tc_kernel|sample_frames_persistent' --launch-skip 9 -c 10 -o gpurun_out/r02_new_kernels -f python scripts/prof_trunk.py fp16 256 2 > gpurun_out/ncu_full_new.log 2>&1 ; echo "ncu full rc=$?" ; ls -la gpurun_out/*.ncu-rep | tail -3
