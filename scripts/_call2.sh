set -u
mkdir -p gpurun_out
AB_CFGS='SERL_PDL=0 SERL_PDL=1 SERL_CAM_STREAMS=1 SERL_PDL=1,SERL_CAM_STREAMS=1'
i=0
for cfg in $AB_CFGS; do
  i=$((i+1)); cfg=${cfg//,/ }
  env $cfg SERL_BENCH_SKIP_SINGLE=1 SERL_BENCH_SKIP_CPU=1 timeout 600 python bench.py --steps 50 --warmup 3 > gpurun_out/bench2_ab_$i.log 2> gpurun_out/bench2_ab_$i.err
  echo "[$cfg] rc=$? $(python -c "
import json
try:
    d=[json.loads(l) for l in open('gpurun_out/bench2_ab_$i.log') if l.startswith('{')][-1]
    print('value %.1f sus %.1f e2e %.1f trunk_ms %.3f frac %.3f samp_frac %.3f sections %s' % (d['value'], d['sustained']['value'], d['e2e']['value'], d['roofline']['ms_per_step'], d['roofline']['frac'], d['sampler']['frac'], {k: v for k, v in d['sections_ms'].items() if k != 'note'}))
except Exception as e:
    print('no line', e)
")"
done
echo "== trunk only"
python scripts/prof_trunk.py fp16 256 5 2>&1 | tail -1
echo "== ncu launch list"
SERL_BENCH_SKIP_SINGLE=1 SERL_BENCH_SKIP_CPU=1 timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 1100 --csv --log-file gpurun_out/launches_r02a.csv python bench.py --steps 2 --warmup 1 --sustain-s 0 > gpurun_out/ncu_bench_r02a.log 2>&1 ; echo "ncu rc=$?"
echo "== ncu --set full (trunk kernels)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'conv3x3_res_kernel|conv3x3s2_res_kernel|stem2_tc_kernel|pool_finish|stem_prep' --launch-skip 11 -c 11 -o gpurun_out/r02a_trunk -f python scripts/prof_trunk.py fp16 256 1 > gpurun_out/ncu_full_r02a.log 2>&1 ; echo "ncu full rc=$?" ; ls -la gpurun_out/*.ncu-rep | tail -3
