#!/bin/bash
# Multi-GPU round trip (run with `gpurun --gpus N`): the 2-GPU hardware DP-parity test, then the bench at every power of two up
# to the GPUs on the box - strong scaling of the headline configuration (global batch 256) and, on 8 GPUs, BASELINE configs[3]
# (global batch 2048).  Output in gpurun_out/scale_*.log.
set -u
mkdir -p gpurun_out
NG=$(nvidia-smi -L | wc -l)
echo "GPUs: $NG"
ENVS="${ENVS:-}"
if [ "$NG" -ge 2 ]; then
  echo "== DP parity (2 GPUs)"; timeout 600 python -m pytest tests/test_dp_gpu.py -m gpu -q -s --timeout=500 -p no:cacheprovider > gpurun_out/pytest_dp.log 2>&1; echo "dp rc=$?"; tail -4 gpurun_out/pytest_dp.log
fi
summ() { python -c "
import json,sys
try:
    d=[json.loads(l) for l in open('$1') if l.startswith('{')][-1]
    print('n=%d value %.1f e2e %.1f ms %.3f launches %d identical %s sections %s' % (d['n_gpus'], d['value'], d['e2e']['value'], d['ms_per_step'], d['gpu_launches'], d['replicas_identical'], {k: v for k, v in d['sections_ms'].items() if k != 'note'}))
except Exception as e:
    print('no line', e)
"; }
for n in 1 2 4 8; do
  [ "$n" -gt "$NG" ] && break
  if [ "$n" -eq 1 ]; then
    env $ENVS SERL_BENCH_SKIP_SINGLE=1 SERL_BENCH_SKIP_CPU=1 timeout 600 python bench.py --gpus 1 --steps 100 --warmup 5 > gpurun_out/scale_n1.log 2> gpurun_out/scale_n1.err
  else
    env $ENVS SERL_BENCH_SKIP_CPU=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600 + n)) bench.py --gpus $n --steps 100 --warmup 5 > gpurun_out/scale_n$n.log 2> gpurun_out/scale_n$n.err
  fi
  echo "[N=$n] rc=$? $(summ gpurun_out/scale_n$n.log)"; tail -2 gpurun_out/scale_n$n.err
done
if [ "$NG" -ge 8 ]; then
  env $ENVS SERL_BENCH_SKIP_CPU=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29700 bench.py --gpus 8 --steps 50 --warmup 5 --batch 2048 > gpurun_out/scale_b2048_n8.log 2> gpurun_out/scale_b2048_n8.err
  echo "[configs[3]: B=2048, N=8] rc=$? $(summ gpurun_out/scale_b2048_n8.log)"; tail -2 gpurun_out/scale_b2048_n8.err
fi
