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
SERL_BENCH_SKIP_CPU=1 timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29608 bench.py --gpus 8 --steps 100 --warmup 5 > gpurun_out/scale_n8_pipe.log 2> gpurun_out/scale_n8_pipe.err
echo "[N=8 B=256 pipeline] rc=$? $(summ gpurun_out/scale_n8_pipe.log)"; grep -v "OMP_NUM\|\*\*\*" gpurun_out/scale_n8_pipe.err | tail -3
SERL_BENCH_SKIP_CPU=1 timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29604 bench.py --gpus 4 --steps 100 --warmup 5 > gpurun_out/scale_n4_pipe.log 2> gpurun_out/scale_n4_pipe.err
echo "[N=4 B=256 pipeline] rc=$? $(summ gpurun_out/scale_n4_pipe.log)"; grep -v "OMP_NUM\|\*\*\*" gpurun_out/scale_n4_pipe.err | tail -3
