set -u
mkdir -p gpurun_out
summ() { python -c "
import json,sys
try:
    d=[json.loads(l) for l in open('$1') if l.startswith('{')][-1]
    print('n=%d value %.1f sus %.1f e2e %.1f ms %.3f launches %d identical %s sections %s' % (d['n_gpus'], d['value'], d['sustained']['value'], d['e2e']['value'], d['ms_per_step'], d['gpu_launches'], d['replicas_identical'], {k: v for k, v in d['sections_ms'].items() if k != 'note'}))
except Exception as e:
    print('no line', e)
"; }
for pl in 1 0; do
SERL_PIPELINE=$pl SERL_BENCH_SKIP_CPU=1 timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2960$pl bench.py --gpus 2 --steps 100 --warmup 5 > gpurun_out/scale_n2_p$pl.log 2> gpurun_out/scale_n2_p$pl.err
echo "[N=2 pipeline=$pl] rc=$? $(summ gpurun_out/scale_n2_p$pl.log)"; grep -v "OMP_NUM\|\*\*\*" gpurun_out/scale_n2_p$pl.err | tail -3
done
