"""CPU, world_size 2, gloo: the data-parallel host logic of `SACAgent._allreduce` - which gradient segment is exchanged
for a critic vs an actor/temperature step, mean semantics (kernels pre-scale by 1/world, the collective sums), info
averaging, and that both ranks end with identical buffers.  Kernels are replaced by a recorder (dry run)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from serl_b200 import _lib as L
    real = L.call
    scales = []

    def fake(name, *a):
        if name.startswith("serl_host_"):
            return real(name, *a)
        if name == "serl_critic_loss":
            scales.append(("critic", float(a[10])))          # grad_scale argument
        if name == "serl_actor_loss":
            scales.append(("actor", float(a[12])))
        return 0

    class Ev:
        def record(self): pass
        def synchronize(self): pass
        def make_current_stream_wait(self): pass

    L.call, L.require_cuda, L.stream_ptr, L.new_event, L.pin = fake, (lambda d: None), (lambda: 0), (lambda: Ev()), (lambda t: t)
    from helpers import fake_env, random_transitions
    from serl_b200.utils.launcher import make_drq_agent, make_replay_buffer
    cams = ("front",)
    rb = make_replay_buffer(fake_env(cams, 128), capacity=40, type="memory_efficient_replay_buffer", image_keys=list(cams), device="cpu",
                            seed=100 + rank)
    trs = random_transitions(np.random.default_rng(rank), 30, cams, 128)
    for tr in trs:
        rb.insert(tr)
    agent = make_drq_agent(7, trs[0]["observations"], trs[0]["actions"], image_keys=cams, encoder_type="resnet-pretrained", device="cpu")
    agent.data_parallel = True
    agent.use_cuda_graphs = False
    st = agent._store
    eng = agent._engine(4)
    # rank-specific gradients / infos stand in for what the (no-op) kernels would have written
    st.grad.copy_(torch.arange(st.n, dtype=torch.float32) * (rank + 1))
    eng.info[:12] = float(rank + 1)
    agent.update_critics(rb.sample(4, pack_obs_and_next_obs=True))
    g_after_critic = st.grad.clone()
    info_after = eng.info[:12].clone()
    st.grad.copy_(torch.arange(st.n, dtype=torch.float32) * (rank + 1))
    agent.update_high_utd(rb.sample(4, pack_obs_and_next_obs=True), utd_ratio=1)
    torch.save(dict(seg=st.seg_end, n=st.n, g_critic=g_after_critic, g_utd=st.grad.clone(), info=info_after, scales=scales), out.format(rank))
    dist.destroy_process_group()


def test_allreduce_segments_and_mean(tmp_path):
    world, port = 2, 29000 + os.getpid() % 2000
    out = str(tmp_path / "rank{}.pt")
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    r0, r1 = torch.load(out.format(0)), torch.load(out.format(1))
    seg, n = r0["seg"], r0["n"]
    base = torch.arange(n, dtype=torch.float32)
    # critic step: only the critic-tx segment [0, seg0) is exchanged: sum over ranks of base*(rank+1) = 3*base
    for r, k in ((r0, 1.0), (r1, 2.0)):
        torch.testing.assert_close(r["g_critic"][:seg[0]], 3.0 * base[:seg[0]])
        torch.testing.assert_close(r["g_critic"][seg[0]:], k * base[seg[0]:])        # untouched elsewhere
    torch.testing.assert_close(r0["g_critic"][:seg[0]], r1["g_critic"][:seg[0]], rtol=0, atol=0)   # replicas identical
    # update_high_utd: critic segment again, then the actor + temperature segments
    torch.testing.assert_close(r0["g_utd"], 3.0 * base)
    torch.testing.assert_close(r0["g_utd"], r1["g_utd"], rtol=0, atol=0)
    # infos are averaged (pmean of aux); kernels were asked to pre-scale gradients by 1/world
    torch.testing.assert_close(r0["info"], torch.full((12,), 1.5))
    assert r0["scales"] and all(abs(v - 0.5) < 1e-12 for _, v in r0["scales"])
    assert {k for k, _ in r0["scales"]} == {"critic", "actor"}
