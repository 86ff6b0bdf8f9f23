set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_tgemm_gpu.py tests/test_fused_heads_gpu.py tests/test_b256_fp16_gpu.py -m gpu -q -x --timeout=600 -p no:cacheprovider > gpurun_out/pytest_fused.log 2>&1; echo "fused rc=$?"; grep -E "passed|failed|Error|error|assert" gpurun_out/pytest_fused.log | tail -10
timeout 300 python scripts/bench_tgemm.py 2>&1 | tail -6
SERL_BENCH_SKIP_SINGLE=1 SERL_BENCH_SKIP_CPU=1 timeout 600 python bench.py --steps 50 --warmup 3 > gpurun_out/bench6.log 2> gpurun_out/bench6.err
echo "rc=$? $(python -c "
import json
try:
    d=[json.loads(l) for l in open('gpurun_out/bench6.log') if l.startswith('{')][-1]
    print('value %.1f sus %.1f e2e %.1f launches %d trunk_ms %.3f frac %.3f sections %s' % (d['value'], d['sustained']['value'], d['e2e']['value'], d['gpu_launches'], d['roofline']['ms_per_step'], d['roofline']['frac'], {k: v for k, v in d['sections_ms'].items() if k != 'note'}))
except Exception as e:
    print('no line', e)
")"; tail -3 gpurun_out/bench6.err
SERL_BENCH_SKIP_SINGLE=1 SERL_BENCH_SKIP_CPU=1 timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r02c.csv python bench.py --steps 2 --warmup 1 --sustain-s 0 > gpurun_out/ncu_bench_r02c.log 2>&1 ; echo "ncu rc=$?"
