"""GPU: the fused critic step of the 16-bit builds (serl_b200/heads_fused.py: TF32 tensor-core GEMMs with TMA-fed operands and
LayerNorm / head epilogues, batched problems) against (a) the float64 oracle and (b) the per-op chain it replaces
(SERL_FUSED_HEADS=0: 3xTF32 GEMMs + separate LayerNorm / reduction kernels) on the same state and batch.
Bars: north_star's 1e-2 for the 16-bit builds on Q / TD target / loss; every gradient leaf within 2e-2 of its own max of the
oracle's (fp16 trunk + TF32 heads) and within 1e-2 of the unfused chain's (TF32 vs 3xTF32 heads, same trunk)."""
import numpy as np
import pytest
import torch

from helpers import fake_env, oracle_cfg_from_agent, oracle_state_from_agent, random_transitions, rel_err, to_numpy_tree

pytestmark = pytest.mark.gpu


def _setup(cams, seed=42, cap=200, n_fill=260):
    from serl_b200.utils.launcher import make_drq_agent, make_replay_buffer
    env = fake_env(cams)
    rb = make_replay_buffer(env, capacity=cap, type="memory_efficient_replay_buffer", image_keys=list(cams), seed=3)
    trs = random_transitions(np.random.default_rng(seed), n_fill, cams)
    for tr in trs:
        rb.insert(tr)
    agent = make_drq_agent(seed, trs[0]["observations"], trs[0]["actions"], image_keys=cams, encoder_type="resnet-pretrained", precision="fp16")
    g = torch.Generator(device="cuda").manual_seed(0)             # biases / LayerNorm offsets off their zero init: every gradient path live
    st = agent._store
    st.params.add_(torch.randn(st.n, device="cuda", generator=g) * 0.05)
    st.target.copy_(st.params + torch.randn(st.n, device="cuda", generator=g) * 0.005)
    st.version += 1
    lam = st.leaf["modules_temperature/lagrange"].offset
    st.params[lam] = -4.0
    st.target[lam] = -4.0
    return agent, rb


@pytest.mark.parametrize("cams,B", [(("front", "wrist"), 12), (("front",), 160)])
def test_fused_critic_step_matches_oracle_and_unfused_chain(cams, B, monkeypatch):
    from oracle import drq as O
    from oracle.replay import unpack
    agent, rb = _setup(cams)
    monkeypatch.setenv("SERL_FUSED_HEADS", "0")
    ref_agent, _ = _setup(cams)
    ref_agent._engine(B)                                              # the engine (and its head path) is built on first use
    monkeypatch.delenv("SERL_FUSED_HEADS")
    ocfg = oracle_cfg_from_agent(agent)
    it = rb.get_iterator(sample_args={"batch_size": B, "pack_obs_and_next_obs": True})
    for step in range(3):                                            # eager, capture + replay, replay
        ostate = oracle_state_from_agent(agent) if B <= 16 else None
        ref_agent._store.params.copy_(agent._store.params); ref_agent._store.target.copy_(agent._store.target)
        ref_agent._store.m.copy_(agent._store.m); ref_agent._store.v.copy_(agent._store.v); ref_agent._store.counts.copy_(agent._store.counts)
        ref_agent.state._rng.copy_(agent.state._rng)
        batch = next(it)
        bd = {k: v for k, v in batch.to_dict().items() if k != "_indices"}
        agent, info = agent.update_critics(batch)
        eng = agent._engines[B]
        assert eng.fused is not None
        ref_agent, rinfo = ref_agent.update_critics(bd)               # same rows as an explicit dict batch (no augmentation difference: same keys)
        reng = ref_agent._engines[B]
        assert reng.fused is None
        for cam in cams:
            assert torch.equal(eng.pix[cam], reng.pix[cam])
        assert rel_err(eng.q.cpu().numpy(), reng.q.cpu().numpy().astype(np.float64)) < 5e-3
        assert rel_err(eng.target_q.cpu().numpy(), reng.target_q.cpu().numpy().astype(np.float64)) < 5e-3
        st, rst = agent._store, ref_agent._store
        worst = ("", 0.0)
        for leaf in st.spec:
            if leaf.group != 0:
                continue
            got, ref = st.view(st.grad, leaf.path).cpu().numpy(), rst.view(rst.grad, leaf.path).cpu().numpy().astype(np.float64)
            e = np.abs(got - ref).max() / max(np.abs(ref).max(), 1e-12)
            worst = max(worst, (leaf.path, e), key=lambda t: t[1])
            assert e < 1e-2, (leaf.path, e)
        print(f"step {step}: fused vs per-op chain: worst gradient leaf {worst[0]} {worst[1]:.2e}")
        if ostate is not None:
            host = unpack(to_numpy_tree(bd))
            oinfo = O.update_critics(ostate, ocfg, host)
            assert rel_err(eng.q.cpu().numpy(), oinfo["critic"]["_q"].numpy()) < 1e-2
            assert rel_err(eng.target_q.cpu().numpy(), oinfo["critic"]["_target_q"].numpy()) < 1e-2
            assert abs(float(info["critic"]["critic_loss"]) - oinfo["critic"]["critic_loss"]) < 1e-2 * max(abs(oinfo["critic"]["critic_loss"]), 1e-6)
            if step == 0:
                for leaf in st.spec:
                    if leaf.group != 0:
                        continue
                    ref = oinfo["_grads"]["critic"][leaf.path].numpy()
                    got = st.view(st.grad, leaf.path).cpu().numpy()
                    assert np.abs(got - ref).max() <= 2e-2 * max(np.abs(ref).max(), 1e-8), leaf.path
    agent.check_status()


def test_learner_iteration_on_the_fp16_build_mixes_fused_critic_and_per_op_actor_step():
    """examples/async_drq_sim/async_drq_sim.py:266-292 on the 16-bit build: `update_critics` (fused head kernels) then
    `update_high_utd(utd_ratio=1)` (critic step on the fused kernels, actor / temperature step on the per-op chain, shared buffers),
    eager then graph-replayed, against the oracle: every info scalar within 1e-2, the key chain bit-exact."""
    from oracle import drq as O
    from oracle.replay import unpack
    cams, B = ("front", "wrist"), 12
    agent, rb = _setup(cams)
    ocfg = oracle_cfg_from_agent(agent)
    it = rb.get_iterator(sample_args={"batch_size": B, "pack_obs_and_next_obs": True})
    close = lambda got, ref: abs(float(got) - ref) <= 1e-2 * max(abs(ref), 1e-2)
    for rep in range(3):
        ostate = oracle_state_from_agent(agent)
        batch = next(it)
        host = unpack(to_numpy_tree({k: v for k, v in batch.to_dict().items() if k != "_indices"}))
        agent, info = agent.update_critics(batch)
        oc = O.update_critics(ostate, ocfg, host)
        assert close(info["critic"]["critic_loss"], oc["critic"]["critic_loss"]), (rep, float(info["critic"]["critic_loss"]), oc["critic"]["critic_loss"])
        np.testing.assert_array_equal(agent.state.rng, ostate.rng)
        ostate = oracle_state_from_agent(agent)
        batch = next(it)
        host = unpack(to_numpy_tree({k: v for k, v in batch.to_dict().items() if k != "_indices"}))
        agent, info = agent.update_high_utd(batch, utd_ratio=1)
        oi = O.update_high_utd(ostate, ocfg, host, 1)
        for k in ("critic_loss", "predicted_qs", "target_qs"):
            assert close(info["critic"][k], oi["critic"][k]), (rep, k, float(info["critic"][k]), oi["critic"][k])
        for k in ("actor_loss", "temperature", "entropy"):
            assert close(info["actor"][k], oi["actor"][k]), (rep, k, float(info["actor"][k]), oi["actor"][k])
        assert close(info["temperature"]["temperature_loss"], oi["temperature"]["temperature_loss"])
        np.testing.assert_array_equal(agent.state.rng, ostate.rng)
    agent.check_status()


def test_fused_actor_temperature_step_matches_per_op_chain(monkeypatch):
    """update(networks_to_update={actor, temperature}) on the fused forward kernels vs SERL_FUSED_ACTOR=0 (per-op chain) from the same
    state, batch and keys: actions, log-probs, Q, every actor / temperature gradient leaf and the proprio encoder's actor-tx twin."""
    cams, B = ("front", "wrist"), 24
    agent, rb = _setup(cams)
    ref_agent, _ = _setup(cams)
    it = rb.get_iterator(sample_args={"batch_size": B, "pack_obs_and_next_obs": True})
    nets = frozenset({"actor", "temperature"})
    for step in range(2):
        for a, b in ((ref_agent._store.params, agent._store.params), (ref_agent._store.target, agent._store.target), (ref_agent._store.m, agent._store.m),
                     (ref_agent._store.v, agent._store.v), (ref_agent._store.counts, agent._store.counts), (ref_agent.state._rng, agent.state._rng)):
            a.copy_(b)
        bd = {k: v for k, v in next(it).to_dict().items() if k != "_indices"}
        agent, info = agent.update(bd, networks_to_update=nets)
        monkeypatch.setenv("SERL_FUSED_ACTOR", "0")
        ref_agent, rinfo = ref_agent.update(bd, networks_to_update=nets)
        monkeypatch.delenv("SERL_FUSED_ACTOR")
        eng, reng = agent._engines[B], ref_agent._engines[B]
        F = eng.F
        assert rel_err(eng.Xc[:, F:].cpu().numpy(), reng.Xc[:, F:].cpu().numpy().astype(np.float64)) < 5e-3          # pi(s)
        # (the per-op chain reuses one log-prob buffer: after the call it holds the temperature pass's, the fused path keeps both)
        assert rel_err(eng.fused.logp_t.cpu().numpy(), reng.logp.cpu().numpy().astype(np.float64)) < 5e-3
        assert rel_err(eng.q.cpu().numpy(), reng.q.cpu().numpy().astype(np.float64)) < 5e-3
        for k in ("actor_loss", "temperature", "entropy"):
            assert abs(float(info["actor"][k]) - float(rinfo["actor"][k])) <= 5e-3 * max(abs(float(rinfo["actor"][k])), 1e-2), k
        assert abs(float(info["temperature"]["temperature_loss"]) - float(rinfo["temperature"]["temperature_loss"])) <= 5e-3 * max(abs(float(rinfo["temperature"]["temperature_loss"])), 1e-3)
        st, rst = agent._store, ref_agent._store
        for leaf in st.spec:
            if leaf.group == 0:
                continue
            got, ref = st.view(st.grad, leaf.path).cpu().numpy(), rst.view(rst.grad, leaf.path).cpu().numpy().astype(np.float64)
            assert np.abs(got - ref).max() <= 1e-2 * max(np.abs(ref).max(), 1e-12), leaf.path
        for path in ("modules_actor/encoder/Dense_0/kernel", "modules_actor/encoder/LayerNorm_0/scale"):
            got, ref = st.aux_view(st.grad, path).cpu().numpy(), rst.aux_view(rst.grad, path).cpu().numpy().astype(np.float64)
            assert np.abs(ref).max() > 0 and np.abs(got - ref).max() <= 1e-2 * np.abs(ref).max(), path
    agent.check_status()
