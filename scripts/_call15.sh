set -u
mkdir -p gpurun_out
for cfg in SERL_TRUNK_SM_LIMIT=112 SERL_TRUNK_SM_LIMIT=96 SERL_TRUNK_SM_LIMIT=112,SERL_PIPELINE=0; do
cfg=${cfg//,/ }
env $cfg SERL_BENCH_SKIP_SINGLE=1 SERL_BENCH_SKIP_CPU=1 timeout 600 python bench.py --steps 100 --warmup 5 > gpurun_out/bench15.log 2> gpurun_out/bench15.err
echo "[$cfg] rc=$? $(python -c "
import json
try:
    d=[json.loads(l) for l in open('gpurun_out/bench15.log') if l.startswith('{')][-1]
    print('value %.1f sus %.1f e2e %.1f launches %d trunk_ms %.3f' % (d['value'], d['sustained']['value'], d['e2e']['value'], d['gpu_launches'], d['roofline']['ms_per_step']))
except Exception as e:
    print('no line', e)
")"; tail -3 gpurun_out/bench15.err
done
