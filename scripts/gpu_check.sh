#!/bin/bash
# One GPU-box round trip: parity tests, smoke, short bench, ncu launch list (+ DRAM bytes).  Everything lands in gpurun_out/.
# (Written before the round-2 kernels became the default: SERL_STEM_V2 / SERL_RES_CONV / SERL_RES_S2 / SERL_CAM_STREAMS are now ON unless set to 0,
# so the "switched on" sections below repeat the default path; scripts/gpu_final.sh is the round-end script.)
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
nproc > gpurun_out/nproc.txt
if [ "${PYTEST:-1}" = "1" ]; then
echo "== pytest gpu" ; timeout 1500 python -m pytest tests -m gpu -q --timeout=600 -p no:cacheprovider ${PYTEST_ARGS:-} > gpurun_out/pytest_gpu.log 2>&1 ; echo "pytest rc=$?" ; tail -25 gpurun_out/pytest_gpu.log
echo "== new kernels switched on: stem v2 / fused trunk / camera streams / PDL through the trunk + agent tests"
SERL_STEM_V2=1 timeout 900 python -m pytest tests/test_trunk_bf16_gpu.py -m gpu -q --timeout=600 -p no:cacheprovider -k "stem or trunk" > gpurun_out/pytest_stem2.log 2>&1 ; echo "stem2 rc=$?" ; tail -6 gpurun_out/pytest_stem2.log
SERL_STEM_V2=1 SERL_RES_CONV=1 SERL_RES_S2=1 SERL_CAM_STREAMS=1 SERL_PDL=1 timeout 900 python -m pytest tests/test_b256_fp16_gpu.py tests/test_agent_gpu.py -m gpu -q --timeout=600 -p no:cacheprovider > gpurun_out/pytest_allnew.log 2>&1 ; echo "allnew rc=$?" ; tail -6 gpurun_out/pytest_allnew.log
SERL_SAMPLER_PERSISTENT=1 timeout 900 python -m pytest tests/test_replay_device.py tests/test_agent_gpu.py -m gpu -q --timeout=600 -p no:cacheprovider > gpurun_out/pytest_psampler.log 2>&1 ; echo "persistent sampler rc=$?" ; tail -4 gpurun_out/pytest_psampler.log
echo "== smoke" ; timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1 ; echo "smoke rc=$?" ; tail -5 gpurun_out/smoke.log
fi
if [ "${BENCH:-1}" = "1" ]; then
echo "== bench" ; timeout 900 python bench.py --steps ${BENCH_STEPS:-50} --warmup 3 ${BENCH_ARGS:-} > gpurun_out/bench.log 2> gpurun_out/bench.err ; echo "bench rc=$?" ; tail -c 3000 gpurun_out/bench.log ; tail -5 gpurun_out/bench.err
fi
if [ "${NCU:-1}" = "1" ]; then
  echo "== ncu launch list"
  SERL_BENCH_SKIP_SINGLE=1 SERL_BENCH_SKIP_CPU=1 timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c ${NCU_COUNT:-1100} --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --sustain-s 0 ${BENCH_ARGS:-} > gpurun_out/ncu_bench.log 2>&1 ; echo "ncu rc=$?"
fi
if [ "${AB:-1}" = "1" ]; then
  echo "== A/B bench runs (50 steps, no single-camera / CPU legs)"
  i=0
  for cfg in ${AB_CFGS:-"SERL_STEM_V2=1" "SERL_STEM_V2=1 SERL_RES_CONV=1" "SERL_STEM_V2=1 SERL_RES_CONV=1 SERL_RES_S2=1" "SERL_SAMPLER_PERSISTENT=1" "SERL_STEM_V2=1 SERL_RES_CONV=1 SERL_RES_S2=1 SERL_PDL=1 SERL_CAM_STREAMS=1 SERL_SAMPLER_PERSISTENT=1"}; do
    i=$((i+1))
    env $cfg SERL_BENCH_SKIP_SINGLE=1 SERL_BENCH_SKIP_CPU=1 timeout 600 python bench.py --steps 50 --warmup 3 > gpurun_out/bench_ab_$i.log 2> gpurun_out/bench_ab_$i.err
    echo "[$cfg] rc=$? $(python -c "
import json,sys
try:
    d=[json.loads(l) for l in open('gpurun_out/bench_ab_$i.log') if l.startswith('{')][-1]
    print('value %.1f e2e %.1f trunk_ms %.3f frac %.3f samp_frac %.3f sections %s' % (d['value'], d['e2e']['value'], d['roofline']['ms_per_step'], d['roofline']['frac'], d['sampler']['frac'], {k: v for k, v in d['sections_ms'].items() if k != 'note'}))
except Exception as e:
    print('no line', e)
")"
    tail -2 gpurun_out/bench_ab_$i.err
  done
fi
