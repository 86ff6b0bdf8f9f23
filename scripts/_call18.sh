set -u
mkdir -p gpurun_out
summ() { python -c "
import json,sys
try:
    d=[json.loads(l) for l in open('$1') if l.startswith('{')][-1]
    print('n=%d value %.1f sus %.1f e2e %.1f ms %.3f launches %d identical %s' % (d['n_gpus'], d['value'], d['sustained']['value'], d['e2e']['value'], d['ms_per_step'], d['gpu_launches'], d['replicas_identical']))
except Exception as e:
    print('no line', e)
"; }
timeout 400 python -m pytest tests/test_dp_gpu.py -m gpu -q --timeout=300 -p no:cacheprovider 2>&1 | tail -2
for sp in 1 0; do
SERL_SPLIT_ALLREDUCE=$sp SERL_BENCH_SKIP_CPU=1 timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2961$sp bench.py --gpus 2 --steps 100 --warmup 5 > gpurun_out/scale_n2_s$sp.log 2> gpurun_out/scale_n2_s$sp.err
echo "[N=2 split_allreduce=$sp] rc=$? $(summ gpurun_out/scale_n2_s$sp.log)"; grep -v "OMP_NUM\|\*\*\*" gpurun_out/scale_n2_s$sp.err | tail -3
done
