"""API conformance: the learner loop of the reference's example script, transcribed line for line from
/root/reference/examples/async_drq_sim/async_drq_sim.py:229-310 + :337-392 (agent / buffer construction), driven through the
`serl_launcher.*` import paths the script uses (served by the shim package at the repo root), with agentlace's TrainerServer,
wandb and jax replaced by stubs.  The jax stub implements just the pytree protocol the script relies on
(`jax.tree_map(jnp.array, agent)`, `jax.device_put(agent, sharding)`, `jax.block_until_ready(agent)`), so the test also pins
that the agent travels through those calls as a leaf-less pytree.

CPU flavour: kernels replaced by a recorder (host logic only).  GPU flavour: the real thing on cuda:0."""
import pickle
import sys
import types

import numpy as np
import pytest
import torch

from helpers import fake_env, random_transitions


class _FakeJax(types.ModuleType):
    """Minimal pytree semantics: registered nodes are flattened by their registered functions, everything else is a leaf."""

    def __init__(self):
        super().__init__("jax")
        self._reg = {}
        self.tree_util = types.SimpleNamespace(register_pytree_node=self._register)
        self.numpy = types.ModuleType("jax.numpy")
        self.numpy.array = np.asarray
        self.leaves_seen = []

    def _register(self, cls, flatten, unflatten):
        if cls in self._reg:
            raise ValueError("duplicate registration")
        self._reg[cls] = (flatten, unflatten)

    def tree_map(self, f, tree):
        if type(tree) in self._reg:
            flatten, unflatten = self._reg[type(tree)]
            children, aux = flatten(tree)
            return unflatten(aux, [self.tree_map(f, c) for c in children])
        self.leaves_seen.append(tree)
        return f(tree)

    def device_put(self, tree, device=None):
        return self.tree_map(lambda x: x, tree)

    def block_until_ready(self, tree):
        return self.tree_map(lambda x: x, tree)


class _Server:
    def __init__(self):
        self.published = []

    def publish_network(self, params):
        self.published.append(params)


@pytest.fixture()
def fake_jax(monkeypatch):
    fj = _FakeJax()
    monkeypatch.setitem(sys.modules, "jax", fj)
    monkeypatch.setitem(sys.modules, "jax.tree_util", fj.tree_util)
    monkeypatch.setitem(sys.modules, "jax.numpy", fj.numpy)
    from serl_b200.agents.continuous.sac import SACAgent, register_pytree
    from serl_b200.agents.continuous.drq import DrQAgent
    assert register_pytree(SACAgent) and register_pytree(DrQAgent)
    return fj


def _learner(fake_jax, device, tmp_path, max_steps, batch_size=8, critic_actor_ratio=4, steps_per_update=2, log_period=1,
             checkpoint_period=2):
    # ---- imports exactly as the script spells them (async_drq_sim.py:19-34) ----
    import jax
    import jax.numpy as jnp
    from serl_launcher.agents.continuous.drq import DrQAgent
    from serl_launcher.data.data_store import MemoryEfficientReplayBufferDataStore
    from serl_launcher.utils.launcher import make_drq_agent, make_replay_buffer, make_wandb_logger
    from serl_launcher.utils.timer_utils import Timer
    from serl_launcher.utils.train_utils import concat_batches
    from serl_b200.utils import checkpoints                      # stands in for flax.training.checkpoints (same call signature)
    cams = ("front", "wrist")
    env = fake_env(cams)
    trs = random_transitions(np.random.default_rng(0), 60, cams)
    image_keys = list(cams)
    sharding = types.SimpleNamespace(replicate=lambda: device)
    # ---- main(): :337-349 ----
    agent = make_drq_agent(seed=42, sample_obs=trs[0]["observations"], sample_action=trs[0]["actions"], image_keys=image_keys,
                           encoder_type="resnet-pretrained", **({"device": device} if device == "cpu" else {}))
    assert isinstance(agent, DrQAgent)
    before = agent
    agent = jax.device_put(jax.tree_map(jnp.array, agent), sharding.replicate())
    assert agent is before and fake_jax.leaves_seen == []            # leaf-less pytree: nothing was copied or converted
    # ---- :353-392 ----
    kw = {"device": device} if device == "cpu" else {}
    replay_buffer = make_replay_buffer(env, capacity=200, rlds_logger_path=None, type="memory_efficient_replay_buffer",
                                       image_keys=image_keys, **kw)
    assert isinstance(replay_buffer, MemoryEfficientReplayBufferDataStore)
    demo_buffer = make_replay_buffer(env, capacity=200, type="memory_efficient_replay_buffer", image_keys=image_keys,
                                     preload_rlds_path=None, preload_data_transform=lambda data, metadata: data, **kw)
    demo_path = tmp_path / "demo.pkl"
    with open(demo_path, "wb") as f:
        pickle.dump(trs[:25], f)
    with open(demo_path, "rb") as f:                                 # :389-392
        trajs = pickle.load(f)
        for traj in trajs:
            demo_buffer.insert(traj)
    for tr in trs:                                                   # the actor's stream (agentlace server thread calls insert)
        replay_buffer.insert(tr)
    wandb_logger = make_wandb_logger(project="serl_dev", description="conformance", debug=True)
    server = _Server()
    # ---- learner(): :229-310 ----
    update_steps = 0
    server.publish_network(agent.state.params)
    single_buffer_batch_size = batch_size // 2
    demo_iterator = demo_buffer.get_iterator(sample_args={"batch_size": single_buffer_batch_size, "pack_obs_and_next_obs": True},
                                             device=sharding.replicate())
    replay_iterator = replay_buffer.get_iterator(sample_args={"batch_size": single_buffer_batch_size, "pack_obs_and_next_obs": True},
                                                 device=sharding.replicate())
    timer = Timer()
    for step in range(max_steps):
        for critic_step in range(critic_actor_ratio - 1):
            with timer.context("sample_replay_buffer"):
                batch = next(replay_iterator)
                if demo_iterator is not None:
                    demo_batch = next(demo_iterator)
                    batch = concat_batches(batch, demo_batch, axis=0)
            with timer.context("train_critics"):
                agent, critics_info = agent.update_critics(batch)
        with timer.context("train"):
            batch = next(replay_iterator)
            if demo_iterator is not None:
                demo_batch = next(demo_iterator)
                batch = concat_batches(batch, demo_batch, axis=0)
            agent, update_info = agent.update_high_utd(batch, utd_ratio=1)
        if step > 0 and step % steps_per_update == 0:
            agent = jax.block_until_ready(agent)
            server.publish_network(agent.state.params)
        if update_steps % log_period == 0 and wandb_logger:
            wandb_logger.log(update_info, step=update_steps)
            wandb_logger.log({"timer": timer.get_average_times()}, step=update_steps)
        if checkpoint_period and update_steps % checkpoint_period == 0:
            checkpoints.save_checkpoint(str(tmp_path / "ckpt"), agent.state, step=update_steps, keep=20)
        update_steps += 1
    return agent, server, wandb_logger, critics_info, update_info, tmp_path / "ckpt"


def _check_outputs(agent, server, logger, critics_info, update_info, max_steps):
    assert set(critics_info) == {"critic", "critic_lr", "actor_lr", "temperature_lr"}
    assert set(update_info["critic"]) == {"critic_loss", "predicted_qs", "target_qs"}
    assert set(update_info["actor"]) == {"actor_loss", "temperature", "entropy"}
    assert set(update_info["temperature"]) == {"temperature_loss"}
    assert agent.state.step == max_steps * 5                         # 3 update_critics + 2 updates inside update_high_utd per iteration
    assert len(server.published) == 1 + sum(1 for s in range(max_steps) if s > 0 and s % 2 == 0)
    tree = server.published[-1]                                       # Flax-layout params tree for the JAX actor
    assert set(tree) == {"modules_actor", "modules_critic", "modules_temperature"}
    enc = tree["modules_actor"]["encoder"]
    assert {"encoder_front", "encoder_wrist", "Dense_0", "LayerNorm_0"} <= set(enc)
    assert enc["encoder_front"]["pretrained_encoder"]["conv_init"]["kernel"].shape == (7, 7, 3, 64)
    assert tree["modules_critic"]["network"]["Dense_0"]["kernel"].shape == (10, 576 + 4, 256)
    keys = {k for _, rec in logger.history for k in rec}
    assert {"critic/critic_loss", "actor/actor_loss", "temperature/temperature_loss", "critic_lr", "timer/train_critics"} <= keys


def test_learner_loop_host_logic_cpu(fake_jax, tmp_path, monkeypatch):
    from serl_b200 import _lib as L
    real_call = L.call
    monkeypatch.setattr(L, "call", lambda name, *a: real_call(name, *a) if name.startswith("serl_host_") else 0)
    monkeypatch.setattr(L, "require_cuda", lambda d: None)
    monkeypatch.setattr(L, "stream_ptr", lambda: 0)
    ev = types.SimpleNamespace(record=lambda: None, synchronize=lambda: None, make_current_stream_wait=lambda: None)
    monkeypatch.setattr(L, "new_event", lambda: ev)
    monkeypatch.setattr(L, "pin", lambda t: t)
    monkeypatch.setattr(L, "launch_count", lambda: 0)
    out = _learner(fake_jax, "cpu", tmp_path, max_steps=3)
    _check_outputs(*out[:5], max_steps=3)


@pytest.mark.gpu
def test_learner_loop_on_gpu_with_checkpoint_roundtrip(fake_jax, tmp_path):
    """The same loop on cuda:0 (fp32 build, CUDA graphs on), then f2: checkpoint -> fresh agent -> restore must reproduce
    params, target params, all three Adam states (incl. the actor-tx twin of the proprio encoder) and the rng bit for bit,
    and the restored agent's next update must equal the original's."""
    from serl_b200.utils import checkpoints
    from serl_b200.utils.launcher import make_drq_agent
    agent, server, logger, critics_info, update_info, ckpt = _learner(fake_jax, "cuda", tmp_path, max_steps=5)
    _check_outputs(agent, server, logger, critics_info, update_info, max_steps=5)
    for k in ("critic_loss", "predicted_qs"):
        assert np.isfinite(float(update_info["critic"][k]))
    agent.check_status()
    checkpoints.save_checkpoint(str(ckpt), agent.state, step=999, keep=20)
    cams = ("front", "wrist")
    trs = random_transitions(np.random.default_rng(0), 2, cams)
    fresh = make_drq_agent(seed=7, sample_obs=trs[0]["observations"], sample_action=trs[0]["actions"], image_keys=list(cams),
                           encoder_type="resnet-pretrained")
    assert not torch.equal(fresh._store.params, agent._store.params)
    restored = checkpoints.restore_checkpoint(str(ckpt), fresh.state)
    fresh = fresh.replace(state=restored)
    for name in ("params", "target", "m", "v"):
        a, b = getattr(agent._store, name), getattr(fresh._store, name)
        off = agent._store.info_off
        mask = torch.ones_like(a, dtype=torch.bool)
        mask[off:off + 16] = False                                   # info gap holds no parameters
        if name in ("params", "target"):
            mask[agent._store.n_main:] = False                       # params / target have no aux part
        assert torch.equal(a[mask], b[mask]), name
    assert torch.equal(agent._store.counts, fresh._store.counts) and fresh.state.step == agent.state.step
    np.testing.assert_array_equal(agent.state.rng, fresh.state.rng)
    for cam in cams:
        for k, t in agent._trunk[cam].items():
            assert torch.equal(t, fresh._trunk[cam][k]), k
    # same next step from both (dict batch -> same rows; same rng -> same crops / noise)
    rng = np.random.default_rng(5)
    B = 4
    batch = {"observations": {**{c: rng.integers(0, 256, (B, 2, 128, 128, 3), dtype=np.uint8) for c in cams},
                              "state": rng.standard_normal((B, 1, 7)).astype(np.float32)},
             "next_observations": {"state": rng.standard_normal((B, 1, 7)).astype(np.float32)},
             "actions": rng.uniform(-1, 1, (B, 4)).astype(np.float32), "rewards": rng.random(B).astype(np.float32),
             "masks": np.ones(B, np.float32), "dones": np.zeros(B, bool)}
    agent.update_high_utd(batch, utd_ratio=1)
    fresh.update_high_utd(batch, utd_ratio=1)
    assert torch.equal(agent._store.params[:agent._store.n_main], fresh._store.params[:agent._store.n_main])
