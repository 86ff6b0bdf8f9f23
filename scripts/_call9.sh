set -u
mkdir -p gpurun_out
NG=$(nvidia-smi -L | wc -l); echo "GPUs: $NG"
timeout 400 python -m pytest tests/test_dp_gpu.py -m gpu -q --timeout=300 -p no:cacheprovider > gpurun_out/pytest_dp.log 2>&1; echo "dp rc=$?"; tail -3 gpurun_out/pytest_dp.log
summ() { python -c "
import json,sys
try:
    d=[json.loads(l) for l in open('$1') if l.startswith('{')][-1]
    print('n=%d value %.1f sus %.1f e2e %.1f ms %.3f launches %d identical %s trunk_frac %.3f sections %s' % (d['n_gpus'], d['value'], d['sustained']['value'], d['e2e']['value'], d['ms_per_step'], d['gpu_launches'], d['replicas_identical'], d['roofline']['frac'], {k: v for k, v in d['sections_ms'].items() if k != 'note'}))
except Exception as e:
    print('no line', e)
"; }
SERL_BENCH_SKIP_CPU=1 timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29608 bench.py --gpus 8 --steps 100 --warmup 5 > gpurun_out/scale_n8.log 2> gpurun_out/scale_n8.err
echo "[N=8 B=256] rc=$? $(summ gpurun_out/scale_n8.log)"; tail -2 gpurun_out/scale_n8.err
SERL_BENCH_SKIP_CPU=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29700 bench.py --gpus 8 --steps 50 --warmup 5 --batch 2048 > gpurun_out/scale_b2048_n8.log 2> gpurun_out/scale_b2048_n8.err
echo "[configs[3]: B=2048, N=8] rc=$? $(summ gpurun_out/scale_b2048_n8.log)"; tail -2 gpurun_out/scale_b2048_n8.err
