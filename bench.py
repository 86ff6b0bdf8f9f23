#!/usr/bin/env python
"""bench.py - DrQ critic grad-steps/sec on B200 (BASELINE.json metric), one JSON line on stdout.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--precision fp32|bf16]
  torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...        (N > 1)

Workload = BASELINE.json configs[1]: `async_drq_sim`, single 128x128 camera, batch 256, replay 100k in HBM.
A "step" = one critic gradient step (`update_critics` equivalent) INCLUDING replay sampling + DrQ shift
(SURVEY.md §8d unit of work).  N > 1: the global batch of 256 is split across ranks (strong scaling), each rank
owns a replay shard, one gradient all-reduce(mean) per step.

  value      steps/s with everything resident in HBM, CUDA-event timed, max over ranks.
  e2e        the same through the public API with host buffers: every step inserts one fresh transition from
             pinned host memory (H2D), draws the batch with the replay iterator, runs agent.update_critics and
             reads the loss back (D2H).
  roofline   frozen ResNet-10 trunk (the dominant kernels): algorithmic FLOPs / CUDA-event time of the trunk
             section inside the timed steps, vs MEASURED_PEAKS.json's sustained bf16 tensor peak; `sampler` gives
             the HBM roofline of the sampler/crop kernel.
  cpu_baseline / --impl reference   the CPU restatement of the reference step (oracle/, torch-CPU fp32, all host
             threads; jax is not installable in this image) on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

TRUNK_GFLOP_PER_IMAGE = 0.5804          # 290,193,408 MAC
FRAME_BYTES = 128 * 128 * 3


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--precision", default=os.environ.get("SERL_PRECISION", "fp16"), choices=["fp32", "bf16", "fp16"],
                    help="trunk arithmetic: fp16 (default: tensor cores, fp32 accumulate, meets the 1e-2 bar), bf16, or fp32 (1e-5 parity build)")
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--cams", type=int, default=1)
    ap.add_argument("--capacity", type=int, default=100_000)
    ap.add_argument("--ref-rows", type=int, default=16, help="rows of the batch the CPU reference processes per step")
    return ap.parse_args()


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], tensor=d.get("bf16_tflops_sustained", d["bf16_tflops"]), src="measured (MEASURED_PEAKS.json, sustained bf16)")
    return dict(hbm=6650.0, tensor=1400.0, src="fallback (B200_PROFILING.md)")


# ------------------------------------------------------------------------------------------------
# CPU reference arm: the oracle port of the reference step (sample on host + update_critics), bounded sample
# ------------------------------------------------------------------------------------------------
def cpu_reference_steps(args, steps, warmup, rows, budget_s=None):
    import torch
    from helpers import random_transitions
    from oracle import drq as O
    from oracle.replay import OracleFrameRing, unpack
    from serl_b200.params import init_trainable, init_trunk, trainable_spec
    cores = min(os.cpu_count() or 1, 32)                  # beyond ~32 threads the small convs of a bounded sample only contend
    torch.set_num_threads(cores)
    cams = tuple(f"cam{i}" for i in range(args.cams))
    rng = np.random.default_rng(0)
    spec = trainable_spec(cams, 7, 4, 10, True)
    params = {k: torch.as_tensor(v) for k, v in init_trainable(rng, spec, 1e-2).items()}
    for cam in cams:
        for k, v in init_trunk(rng).items():
            params[f"modules_actor/encoder/encoder_{cam}/pretrained_encoder/{k}"] = torch.as_tensor(v)
    state = O.OracleState.create(params, np.array([0, 42], np.uint32), torch.float32)
    cfg = O.OracleConfig(cams=cams)
    ring = OracleFrameRing(1200, cams, (128, 128, 3), 1, 7, 4)
    for tr in random_transitions(rng, 1000, cams, mean_ep=100):
        ring.insert(tr)
    times, t_begin = [], time.perf_counter()
    for s in range(warmup + steps):
        t0 = time.perf_counter()
        _, packed = ring.sample(0, s, rows)
        O.update_critics(state, cfg, unpack(packed), dtype=torch.float32)
        dt = time.perf_counter() - t0
        if s >= warmup:
            times.append(dt)
        if budget_s is not None and times and time.perf_counter() - t_begin > budget_s:
            break                                          # bounded sample: stop once the time budget is spent
    t = sum(times) / len(times)
    # a full step processes `batch` rows; the sample processed `rows`: scale linearly (trunk-dominated, per-row cost)
    return (rows / args.batch) / t, t, cores, len(times)


def run_reference(args):
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    v, t, cores, done = cpu_reference_steps(args, args.steps, min(args.warmup, 1), args.ref_rows, budget_s=150.0)
    sample = (f"{done} timed steps (150 s budget) of {args.ref_rows} of {args.batch} rows per step (host numpy sampling + torch-CPU fp32 restatement of update_critics, "
              f"trunk shared between policy/critic/target like the B200 path; the JAX reference recomputes it 3x); "
              f"steps/s scaled by rows/batch")
    line = {"metric": "drq_critic_grad_steps_per_sec", "value": v, "unit": "steps/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 / v, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "impl": "reference",
            "config": workload_config(args),
            "cpu_baseline": {"value": v, "unit": "steps/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": v, "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


TRUNK_KERNELS = {
    False: "frozen ResNet-10 trunk, fp32 build (conv_igemm_f32 + groupnorm_f32 + maxpool3x3s2_f32)",
    True: "frozen ResNet-10 trunk, tcgen05 build (stem_tc + conv3x3_tc + conv_tc kernels and their elementwise GroupNorm / pool / residual passes)",
}


def trunk_traffic(args):
    """DRAM bytes per step of the trunk kernels (dram__bytes_read.sum + dram__bytes_write.sum summed over the trunk's launches of
    one step) from the committed ncu capture of this same command, or None when no capture matches the configuration."""
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", "trunk_traffic.json")
    try:
        with open(path) as f:
            t = json.load(f)
        key = f"{args.precision}_b{args.batch}_c{args.cams}"
        return t.get(key, {}).get("dram_bytes_per_step")
    except (OSError, ValueError):
        return None


def workload_config(args):
    return {"workload": f"async_drq_sim: {args.cams}x 128x128x3 camera, batch {args.batch} (global), replay {args.capacity} in HBM, "
                        "critic grad step incl. sampling + DrQ shift", "global_batch": args.batch, "cams": args.cams,
            "replay_capacity": args.capacity, "parallelism": f"dp{args.gpus}", "precision": args.precision,
            "arithmetic": ("frozen ResNet-10 trunk: 16-bit operands on tcgen05 tensor cores with fp32 accumulation; trainable heads, losses, "
                           "Adam in fp32" if args.precision != "fp32" else "everything fp32 (CUDA cores): the 1e-5 parity build"),
            "l2": "inputs exceed L2: each step gathers fresh random frames from a multi-GB replay"}


# ------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = "timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.p, self.index = None, index

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50", "-i", str(self.index)],
                                      stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:                       # noqa: BLE001
            self.p = None

    def stop(self, t_begin=None, t_end=None):
        import datetime
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.p.terminate()
        try:
            out, _ = self.p.communicate(timeout=5)
        except Exception:                       # noqa: BLE001
            self.p.kill()
            out = ""
        sm, mx, reasons = [], None, set()
        for ln in out.strip().splitlines():
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 8:
                continue
            try:
                ts = datetime.datetime.strptime(f[0], "%Y/%m/%d %H:%M:%S.%f").timestamp()
                if t_begin is not None and not (t_begin - 0.05 <= ts <= t_end + 0.05):
                    continue
                sm.append(float(f[1])); mx = float(f[2])
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


def fill_ring_synthetic(rb, seed):
    """SURVEY.md §8d synthetic replay: random frames, episodes of 100 (1 filler slot in 101), N(0,1) state, U(-1,1) actions."""
    import torch
    g = torch.Generator(device="cuda").manual_seed(seed)
    cap = rb._capacity
    for c in rb.cams:
        fr = rb.frames[c]
        chunk = 4096
        for lo in range(0, cap, chunk):
            hi = min(cap, lo + chunk)
            fr[lo:hi] = torch.randint(0, 256, (hi - lo, *fr.shape[1:]), dtype=torch.uint8, device="cuda", generator=g)
    slots = torch.arange(cap, device="cuda")
    valid = (slots % 101) != 0
    rb.valid.copy_(valid.to(torch.uint8))
    rb._valid_host[:] = valid.cpu().numpy()
    rb.state.copy_(torch.randn(rb.state.shape, device="cuda", generator=g))
    rb.next_state.copy_(torch.randn(rb.state.shape, device="cuda", generator=g))
    rb.actions.copy_(torch.rand(rb.actions.shape, device="cuda", generator=g) * 2 - 1)
    rb.rewards.copy_(torch.rand(cap, device="cuda", generator=g))
    ends = (slots % 101) == 100
    rb.masks.copy_((~ends).float())
    rb.dones.copy_(ends.to(torch.uint8))
    rb._size, rb._insert_index, rb._first = cap, 0, False
    rb.size_dev.fill_(cap)


def measure_dual_camera_rlpd(args, steps=50):
    """Supplementary measurement on BASELINE configs[2]: dual 128x128 cameras, batch 256 drawn 50/50 from the online ring and a
    20-trajectory demo ring (RLPD), same metric, same timing rules (device-resident value + e2e with host insert / loss readback)."""
    import torch
    from helpers import fake_env, random_transitions
    from serl_b200.utils.launcher import make_drq_agent, make_replay_buffer
    from serl_b200.utils.train_utils import concat_batches
    cams, half = ("cam0", "cam1"), args.batch // 2
    env = fake_env(cams)
    rb = make_replay_buffer(env, capacity=args.capacity, type="memory_efficient_replay_buffer", image_keys=list(cams), seed=2000)
    demo = make_replay_buffer(env, capacity=20 * 101, type="memory_efficient_replay_buffer", image_keys=list(cams), seed=2001)
    fill_ring_synthetic(rb, seed=7)
    fill_ring_synthetic(demo, seed=8)
    rng = np.random.default_rng(1)
    trs = random_transitions(rng, 8, cams, mean_ep=1000)
    agent = make_drq_agent(42, trs[0]["observations"], trs[0]["actions"], image_keys=cams, encoder_type="resnet-pretrained", precision=args.precision)
    it = rb.get_iterator(sample_args={"batch_size": half, "pack_obs_and_next_obs": True})
    dit = demo.get_iterator(sample_args={"batch_size": args.batch - half, "pack_obs_and_next_obs": True})
    nxt = lambda: concat_batches(next(it), next(dit), axis=0)
    for _ in range(5):
        agent.update_critics(nxt())
    torch.cuda.synchronize()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(steps):
        agent.update_critics(nxt())
    t1.record()
    torch.cuda.synchronize()
    ms = t0.elapsed_time(t1) / steps
    h0, e0 = rb.h2d_bytes, time.perf_counter()
    for s in range(steps):
        rb.insert(trs[s % len(trs)])
        _, info = agent.update_critics(nxt())
        loss = float(info["critic"]["critic_loss"])
    torch.cuda.synchronize()
    e2e_ms = (time.perf_counter() - e0) * 1e3 / steps
    assert np.isfinite(loss)
    out = {"workload": f"async_drq_sim + demos (50/50 RLPD): 2x 128x128x3 cameras, batch {args.batch} = {half} online + {args.batch - half} demo, "
                       f"replay {args.capacity} + {20 * 101} in HBM, critic grad step incl. sampling + DrQ shift",
           "value": 1e3 / ms, "unit": "steps/s", "ms_per_step": ms, "steps": steps,
           "e2e": {"value": 1e3 / e2e_ms, "unit": "steps/s", "h2d_bytes_per_step": (rb.h2d_bytes - h0) / steps, "d2h_bytes_per_step": 4.0}}
    agent._graphs.clear()
    del agent, rb, demo
    torch.cuda.synchronize()
    return out


def run_b200(args):
    import torch
    import torch.distributed as dist
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from helpers import fake_env, random_transitions
    from serl_b200.utils.launcher import make_drq_agent, make_replay_buffer
    assert args.batch % world == 0
    B = args.batch // world
    cams = tuple(f"cam{i}" for i in range(args.cams))
    env = fake_env(cams)
    rb = make_replay_buffer(env, capacity=args.capacity // world, type="memory_efficient_replay_buffer", image_keys=list(cams),
                            seed=1000 + rank)                     # rank folded into the sampler stream
    fill_ring_synthetic(rb, seed=rank)
    rng = np.random.default_rng(0)
    sample_tr = random_transitions(rng, 1, cams)[0]
    agent = make_drq_agent(42, sample_tr["observations"], sample_tr["actions"], image_keys=cams, encoder_type="resnet-pretrained",
                           precision=args.precision)
    agent.data_parallel = world > 1
    it = rb.get_iterator(sample_args={"batch_size": B, "pack_obs_and_next_obs": True})
    eng = agent._engine(B)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident throughput (whole step replayed as one CUDA graph) -------------------------------------
    ev = lambda: torch.cuda.Event(enable_timing=True)
    clocks = ClockSampler(local)
    clocks.start()
    for _ in range(max(args.warmup, 3)):
        agent.update_critics(next(it))
    launches0 = agent.kernel_launches
    barrier()
    w0 = time.time()
    t0, t1 = ev(), ev()
    t0.record()
    for _ in range(args.steps):
        agent.update_critics(next(it))
    t1.record()
    barrier()
    clk = clocks.stop(w0, time.time())
    ms = t0.elapsed_time(t1)
    launches = agent.kernel_launches - launches0
    agent.check_status()

    # ---- per-kernel-group durations: the same steps launched eagerly with CUDA events around the sections -----------
    trunk_ev, samp_ev = [], []
    orig_features, orig_load = agent._features, agent._load_batch

    def timed_features(e):
        a, b = ev(), ev(); a.record(); orig_features(e); b.record(); trunk_ev.append((a, b))

    def timed_load(e, batch, **kw):
        a, b = ev(), ev(); a.record(); orig_load(e, batch, **kw); b.record(); samp_ev.append((a, b))

    agent.use_cuda_graphs = False
    agent._features, agent._load_batch = timed_features, timed_load
    for _ in range(min(args.steps, 20)):
        agent.update_critics(next(it))
    barrier()
    agent._features, agent._load_batch = orig_features, orig_load
    del samp_ev
    agent.use_cuda_graphs = True
    trunk_ms = sum(a.elapsed_time(b) for a, b in trunk_ev) / len(trunk_ev)
    # the sampler kernel is ~10x shorter than a host launch: time it as 20 launches captured in one CUDA graph, replayed
    # back to back (each launch draws a fresh batch: the device step counter advances inside the graph)
    handle = next(it)
    reps = 20
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps):
            orig_load(eng, handle, augment=True, keys=agent._keys, graph_mode=True)
    g.replay(); torch.cuda.synchronize()
    a, b = ev(), ev()
    a.record()
    for _ in range(5):
        g.replay()
    b.record(); torch.cuda.synchronize()
    samp_ms = a.elapsed_time(b) / (5 * reps)
    del g

    # ---- end to end through the public API with host buffers --------------------------------------------
    pinned = []
    for tr in random_transitions(rng, 8, cams, mean_ep=1000):
        pinned.append(tr)
    h2d0 = rb.h2d_bytes
    barrier()
    e0 = time.perf_counter()
    d2h = 0
    for s in range(args.steps):
        rb.insert(pinned[s % len(pinned)])                       # fresh transition from host memory -> pinned staging -> HBM
        batch = next(it)
        _, info = agent.update_critics(batch)
        loss = float(info["critic"]["critic_loss"])              # D2H read of the step's result
        d2h += 4
    barrier()
    e2e_s = time.perf_counter() - e0
    h2d = (rb.h2d_bytes - h2d0) / args.steps
    assert np.isfinite(loss)

    replicas_identical = None
    if world > 1:                                   # data-parallel replicas must stay bit-identical (same reduced gradient everywhere)
        mine = agent._store.params.clone()
        ref = mine.clone()
        dist.broadcast(ref, src=0)
        same = torch.tensor([int(torch.equal(mine, ref))], device="cuda")
        dist.all_reduce(same, op=dist.ReduceOp.MIN)
        replicas_identical = bool(same.item())
    tmax = torch.tensor([ms, e2e_s * 1e3], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    ms, e2e_ms = tmax.tolist()
    def shutdown():
        # captured NCCL kernels inside live CUDA graphs can block process-group teardown: drop the graphs first, and never
        # let a teardown problem turn into a hung bench (os._exit after the line is out)
        agent._graphs.clear()
        torch.cuda.synchronize()
        sys.stdout.flush()
        if world > 1:
            os._exit(0)

    if rank != 0:
        shutdown()
        return
    pk = peaks()
    value = args.steps / (ms / 1e3)
    images = 2 * B * args.cams                                    # per rank per step (obs + next_obs, trunk shared)
    trunk_tflops = images * TRUNK_GFLOP_PER_IMAGE / 1e3 / (trunk_ms / 1e3)
    samp_bytes = 2 * B * args.cams * 2 * FRAME_BYTES              # read 2 frames + write 2 crops per sample per camera
    samp_gbs = samp_bytes / 1e9 / (samp_ms / 1e3)
    line = {"metric": "drq_critic_grad_steps_per_sec", "value": value, "unit": "steps/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": {"fp32": "f32", "bf16": "bf16", "fp16": "f16"}[args.precision], "data": "synthetic", "impl": "b200",
            "config": workload_config(args),
            "clocks": clk,
            "e2e": {"value": args.steps / (e2e_ms / 1e3), "unit": "steps/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h / args.steps},
            "gpu_launches": launches, "cuda_graph": True, "replicas_identical": replicas_identical,
            "roofline": {"kernel": TRUNK_KERNELS[args.precision != "fp32"], "bound": "tensor",
                         "achieved": trunk_tflops, "peak": pk["tensor"], "unit": "TFLOP/s", "frac": trunk_tflops / pk["tensor"],
                         "traffic": trunk_traffic(args), "peak_source": pk["src"], "ms_per_step": trunk_ms,
                         "timing": "CUDA events around the trunk section of eagerly launched steps (the headline loop replays a CUDA graph)",
                         "algorithmic": f"{images} images x {TRUNK_GFLOP_PER_IMAGE} GFLOP"},
            "sampler": {"kernel": "sample_frames_kernel", "timing": "20 launches captured in one CUDA graph, replayed 5x, CUDA events", "bound": "hbm", "achieved": samp_gbs, "peak": pk["hbm"], "unit": "GB/s",
                        "frac": samp_gbs / pk["hbm"], "ms_per_step": samp_ms, "algorithmic_bytes": samp_bytes}}
    if world == 1 and args.cams == 1 and not os.environ.get("SERL_BENCH_SKIP_DUAL"):
        # the stock sim script is dual-camera: report BASELINE configs[2] beside the headline configuration
        try:
            line["dual_camera_rlpd"] = measure_dual_camera_rlpd(args)
        except Exception as e:                  # noqa: BLE001
            line["dual_camera_rlpd"] = {"value": None, "error": str(e)}
    try:
        if os.environ.get("SERL_BENCH_SKIP_CPU"):
            raise RuntimeError("skipped (SERL_BENCH_SKIP_CPU)")
        v, t, cores, done = cpu_reference_steps(args, 4, 1, 8, budget_s=20.0)
        line["cpu_baseline"] = {"value": v, "unit": "steps/s", "cores": cores, "kind": "port",
                                "sample": f"{done} timed steps of 8/{args.batch} rows, ~20 s budget (oracle torch-CPU fp32 restatement of sample + update_critics; "
                                          "jax not installable), steps/s scaled by rows/batch"}
    except Exception as e:                      # noqa: BLE001
        line["cpu_baseline"] = {"value": None, "unit": "steps/s", "cores": os.cpu_count(), "kind": "port", "sample": f"failed: {e}"}
    print(json.dumps(line), flush=True)
    shutdown()


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_b200(a)
