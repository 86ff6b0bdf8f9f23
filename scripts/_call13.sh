set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout=900 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; grep -E "passed|failed|Error|error|assert|FAILED" gpurun_out/pytest_gpu.log | tail -20
for cfg in SERL_PIPELINE=1 SERL_PIPELINE=0; do
env $cfg SERL_BENCH_SKIP_SINGLE=1 SERL_BENCH_SKIP_CPU=1 timeout 600 python bench.py --steps 100 --warmup 5 > gpurun_out/bench13.log 2> gpurun_out/bench13.err
echo "[$cfg] rc=$? $(python -c "
import json
try:
    d=[json.loads(l) for l in open('gpurun_out/bench13.log') if l.startswith('{')][-1]
    print('value %.1f sus %.1f e2e %.1f launches %d trunk_ms %.3f frac %.3f samp %.3f sections %s' % (d['value'], d['sustained']['value'], d['e2e']['value'], d['gpu_launches'], d['roofline']['ms_per_step'], d['roofline']['frac'], d['sampler']['frac'], {k: v for k, v in d['sections_ms'].items() if k != 'note'}))
except Exception as e:
    print('no line', e)
")"; tail -3 gpurun_out/bench13.err
done
