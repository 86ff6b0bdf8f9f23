"""GPU: whole DrQ / SAC gradient steps through the serl_launcher-style API vs the float64 oracle.

Tolerances: north_star asks 1e-5 (fp32) on Q-values, losses and sampled actions; parameters after Adam
are compared at 1e-5 relative to the parameter scale; gradients at 1e-4 of their own max (fp32 kernels
against a float64 oracle through ~10 GEMM/norm layers)."""
import numpy as np
import pytest
import torch

from helpers import (fake_env, oracle_cfg_from_agent, oracle_state_from_agent, random_transitions, rel_err, to_numpy_tree)

pytestmark = pytest.mark.gpu
Q_TOL, P_TOL, G_TOL = 1e-5, 1e-5, 2e-4


def _setup(cams, B, seed=42, cap=200, n_fill=260):
    from serl_b200.utils.launcher import make_drq_agent, make_replay_buffer
    env = fake_env(cams)
    rb = make_replay_buffer(env, capacity=cap, type="memory_efficient_replay_buffer", image_keys=list(cams), seed=3)
    rng = np.random.default_rng(seed)
    trs = random_transitions(rng, n_fill, cams)
    for tr in trs:
        rb.insert(tr)
    agent = make_drq_agent(seed, trs[0]["observations"], trs[0]["actions"], image_keys=cams, encoder_type="resnet-pretrained")
    return agent, rb


def _perturb(agent, scale=0.05, seed=0):
    """Move LayerNorm biases / Dense biases off their zero init so every gradient path is exercised."""
    g = torch.Generator(device="cuda").manual_seed(seed)
    st = agent._store
    noise = torch.randn(st.n, device="cuda", generator=g) * scale
    st.params.add_(noise)
    st.target.copy_(st.params + torch.randn(st.n, device="cuda", generator=g) * scale * 0.1)
    st.version += 1
    lam = st.leaf["modules_temperature/lagrange"].offset
    st.params[lam] = -4.0
    st.target[lam] = -4.0


def _flat_grads(info, group_paths):
    return {k: info["_grads"][g][k] for g, paths in group_paths.items() for k in paths}


def _compare_state(agent, ostate, oinfo, what=""):
    """Post-step params / target / rng vs the oracle transition from the SAME pre-step state.

    Adam normalises the step by |g|: an entry whose gradient is at fp32 noise level (|g| << max|g| of its leaf) can
    legitimately move by up to ~2*lr in either direction in float32 (the JAX fp32 reference has the same property
    against a float64 run), so such entries get an lr-sized allowance; well-conditioned entries must match tightly."""
    from serl_b200.params import flatten
    p, tp = flatten(agent.state.params), flatten(agent.state.target_params)
    lr = agent._cfg.lr[0]
    gsum = None
    for g in oinfo.get("_grads_abs_all_calls", oinfo["_grads"]).values():
        gsum = {k: np.abs(v.numpy()) for k, v in g.items()} if gsum is None else {k: gsum[k] + np.abs(g[k].numpy()) for k in gsum}
    for k in p:
        ref, tref = ostate.params[k].numpy(), ostate.target_params[k].numpy()
        scale = max(np.abs(ref).max(), 1e-3)
        g = gsum[k]
        noisy = g < 2e-2 * max(g.max(), 1e-30)
        allow = P_TOL * scale + lr * np.where(noisy, 2.2, 5e-3)
        bad = np.abs(p[k] - ref) > allow
        assert not bad.any(), f"{what}: {k}: {bad.sum()} entries off, worst {np.abs(p[k] - ref).max():.2e} (scale {scale:.2e})"
        bad_t = np.abs(tp[k] - tref) > P_TOL * scale + agent._cfg.tau * allow
        assert not bad_t.any(), f"{what}: target {k}: worst {np.abs(tp[k] - tref).max():.2e}"
    np.testing.assert_array_equal(agent.state.rng, ostate.rng)


@pytest.mark.parametrize("cams,B", [(("front",), 16), (("front", "wrist"), 12)])
def test_update_critics_matches_oracle(cams, B):
    from oracle import drq as O
    from oracle.replay import unpack
    agent, rb = _setup(cams, B)
    _perturb(agent)
    ocfg = oracle_cfg_from_agent(agent)
    it = rb.get_iterator(sample_args={"batch_size": B, "pack_obs_and_next_obs": True})
    for step in range(2):
        ostate = oracle_state_from_agent(agent)
        batch = next(it)
        host = unpack(to_numpy_tree({k: v for k, v in batch.to_dict().items() if k != "_indices"}))
        agent, info = agent.update_critics(batch)
        oinfo = O.update_critics(ostate, ocfg, host)
        eng = agent._engines[B]
        # integer outputs bit-exact: crops applied by the sampler kernel == oracle augmentation
        for cam in cams:
            pix = eng.pix[cam].cpu().numpy()
            np.testing.assert_array_equal(pix[:B], oinfo["_aug"]["observations"][cam][:, 0])
            np.testing.assert_array_equal(pix[B:], oinfo["_aug"]["next_observations"][cam][:, 0])
        assert rel_err(eng.q.cpu().numpy(), oinfo["critic"]["_q"].numpy()) < Q_TOL
        assert rel_err(eng.target_q.cpu().numpy(), oinfo["critic"]["_target_q"].numpy()) < Q_TOL
        for k in ("critic_loss", "predicted_qs", "target_qs"):
            np.testing.assert_allclose(float(info["critic"][k]), oinfo["critic"][k], rtol=Q_TOL, atol=1e-6)
        assert "actor" not in info and "temperature" not in info
        if step == 0:
            st = agent._store
            for leaf in st.spec:
                if leaf.group != 0:
                    continue
                ref = oinfo["_grads"]["critic"][leaf.path].numpy()
                got = st.view(st.grad, leaf.path).cpu().numpy()
                assert np.abs(got - ref).max() <= G_TOL * max(np.abs(ref).max(), 1e-8), leaf.path
        _compare_state(agent, ostate, oinfo, f"step {step}")
    assert agent.state.step == 2
    agent.check_status()


def test_learner_iteration_matches_oracle():
    """One learner iteration of examples/async_drq_sim/async_drq_sim.py:266-292 with critic_actor_ratio=2:
    update_critics, then update_high_utd(utd_ratio=1) - exercises actor + temperature losses and the
    zero-gradient momentum drift of all three Adam txs."""
    from oracle import drq as O
    from oracle.replay import unpack
    cams, B = ("front",), 16
    agent, rb = _setup(cams, B, seed=7)
    _perturb(agent, seed=1)
    ocfg = oracle_cfg_from_agent(agent)
    it = rb.get_iterator(sample_args={"batch_size": B, "pack_obs_and_next_obs": True})
    for rep in range(2):
        ostate = oracle_state_from_agent(agent)
        batch = next(it)
        host = unpack(to_numpy_tree({k: v for k, v in batch.to_dict().items() if k != "_indices"}))
        agent, _ = agent.update_critics(batch)
        oc = O.update_critics(ostate, ocfg, host)
        _compare_state(agent, ostate, oc, f"iteration {rep} critics")
        ostate = oracle_state_from_agent(agent)
        batch = next(it)
        host = unpack(to_numpy_tree({k: v for k, v in batch.to_dict().items() if k != "_indices"}))
        agent, info = agent.update_high_utd(batch, utd_ratio=1)
        oinfo = O.update_high_utd(ostate, ocfg, host, 1)
        for k in ("critic_loss", "predicted_qs", "target_qs"):
            np.testing.assert_allclose(float(info["critic"][k]), oinfo["critic"][k], rtol=Q_TOL, atol=1e-6)
        for k in ("actor_loss", "temperature", "entropy"):
            np.testing.assert_allclose(float(info["actor"][k]), oinfo["actor"][k], rtol=Q_TOL, atol=1e-6)
        np.testing.assert_allclose(float(info["temperature"]["temperature_loss"]), oinfo["temperature"]["temperature_loss"],
                                   rtol=Q_TOL, atol=1e-7)
        if rep == 0:
            st = agent._store
            for leaf in st.spec:
                if leaf.group == 0:
                    continue
                ref = oinfo["_grads"]["actor" if leaf.group == 1 else "temperature"][leaf.path].numpy()
                got = st.view(st.grad, leaf.path).cpu().numpy()
                assert np.abs(got - ref).max() <= G_TOL * max(np.abs(ref).max(), 1e-8), leaf.path
            # the ACTOR loss also differentiates the proprio encoder: Policy's stop_gradient covers the image embeddings
            # only (common/encoding.py:48-49 vs :55-70, sac.py:198-200).  Its gradient lives in the actor-tx twin (aux tail).
            for path in ("modules_actor/encoder/Dense_0/kernel", "modules_actor/encoder/Dense_0/bias",
                         "modules_actor/encoder/LayerNorm_0/scale", "modules_actor/encoder/LayerNorm_0/bias"):
                ref = oinfo["_grads"]["actor"][path].numpy()
                got = st.aux_view(st.grad, path).cpu().numpy()
                assert np.abs(ref).max() > 0, f"oracle: actor loss must reach {path}"
                assert np.abs(got).max() > 0, f"actor-loss gradient into {path} dropped"
                assert np.abs(got - ref).max() <= G_TOL * np.abs(ref).max(), path
            # ... and nothing else under modules_actor/encoder (image heads are behind the stop_gradient)
            for k, v in oinfo["_grads"]["actor"].items():
                if "/encoder_" in k:
                    assert float(v.abs().max()) == 0.0, k
            os_ = agent.state.opt_states
            from serl_b200.params import flatten
            assert np.abs(flatten(os_["actor"]["mu"])["modules_actor/encoder/Dense_0/kernel"]).max() > 0
            assert np.abs(flatten(os_["critic"]["mu"])["modules_actor/encoder/Dense_0/kernel"]).max() > 0
        # update_high_utd = two `update` calls; the oracle's gradient record is of the last one (actor+temperature),
        # the critic-tx entries moved in the first: merge both records for the conditioning mask
        _compare_state(agent, ostate, oinfo, f"iteration {rep}")
        assert float(info["critic_lr"]) == pytest.approx(3e-4)


def test_rlpd_concat_and_dict_batches():
    """50/50 RLPD sampling (async_drq_sim.py:275-277): concat_batches(online, demo) keeps online rows first and the
    crop key index runs over the concatenated batch; a plain dict batch gives the same update as its handle."""
    from oracle import drq as O
    from oracle.replay import concat_batches as oconcat
    from oracle.replay import unpack
    from serl_b200.utils.launcher import make_replay_buffer
    from serl_b200.utils.train_utils import concat_batches
    cams, B = ("front",), 8
    agent, rb = _setup(cams, 2 * B, seed=9)
    demo = make_replay_buffer(fake_env(cams), capacity=60, type="memory_efficient_replay_buffer", image_keys=list(cams), seed=8)
    for tr in random_transitions(np.random.default_rng(5), 50, cams):
        demo.insert(tr)
    _perturb(agent, seed=2)
    ostate, ocfg = oracle_state_from_agent(agent), oracle_cfg_from_agent(agent)
    b1, b2 = rb.sample(B, pack_obs_and_next_obs=True), demo.sample(B, pack_obs_and_next_obs=True)
    both = concat_batches(b1, b2, axis=0)
    h1 = to_numpy_tree({k: v for k, v in b1.to_dict().items() if k != "_indices"})
    h2 = to_numpy_tree({k: v for k, v in b2.to_dict().items() if k != "_indices"})
    host = unpack(oconcat(h1, h2, axis=0))
    agent, info = agent.update_critics(both)
    oinfo = O.update_critics(ostate, ocfg, host)
    np.testing.assert_allclose(float(info["critic"]["critic_loss"]), oinfo["critic"]["critic_loss"], rtol=Q_TOL)
    _compare_state(agent, ostate, oinfo, "rlpd")


def test_dict_batch_equals_handle_batch():
    """A host dict in the reference's packed layout (what its sample() returns) gives bit-identical updates to the lazy handle."""
    cams, B = ("front",), 8
    agent, rb = _setup(cams, B, seed=9)
    agent2, _ = _setup(cams, B, seed=9)
    _perturb(agent, seed=2)
    _perturb(agent2, seed=2)
    torch.testing.assert_close(agent._store.params, agent2._store.params, rtol=0, atol=0)
    b3 = rb.sample(B, pack_obs_and_next_obs=True)
    d3 = to_numpy_tree({k: v for k, v in b3.to_dict().items() if k != "_indices"})
    agent.update_critics(b3)
    agent2.update_critics(d3)
    torch.testing.assert_close(agent._store.params, agent2._store.params, rtol=0, atol=0)
    torch.testing.assert_close(agent._engines[B].q, agent2._engines[B].q, rtol=0, atol=0)
    agent.update_high_utd(b3, utd_ratio=1)
    agent2.update_high_utd(d3, utd_ratio=1)
    torch.testing.assert_close(agent._store.params, agent2._store.params, rtol=0, atol=0)


def oracle_tree(flat):
    from serl_b200.params import nest
    return nest({k: v.numpy() for k, v in flat.items()})


def test_sample_actions_matches_oracle():
    from oracle import drq as O
    from oracle import jax_prng as P
    cams = ("front", "wrist")
    agent, rb = _setup(cams, 4, seed=13)
    _perturb(agent, seed=3)
    ostate, ocfg = oracle_state_from_agent(agent), oracle_cfg_from_agent(agent)
    rng = np.random.default_rng(0)
    obs = {c: rng.integers(0, 256, (1, 128, 128, 3), dtype=np.uint8) for c in cams}
    obs["state"] = rng.standard_normal((1, 7)).astype(np.float32)
    key = P.prng_key(2024)
    a = agent.sample_actions(obs, seed=key)                      # unbatched, like the actor loop (async_drq_sim.py:130-136)
    ob = {k: v[None] for k, v in obs.items()}
    ref = O.sample_actions(ostate, ocfg, ob, seed=key)[0].numpy()
    assert a.shape == (4,) and rel_err(a, ref) < Q_TOL
    am = agent.sample_actions(obs, argmax=True)
    assert rel_err(am, O.sample_actions(ostate, ocfg, ob, argmax=True)[0].numpy()) < Q_TOL
    batch = {c: rng.integers(0, 256, (5, 1, 128, 128, 3), dtype=np.uint8) for c in cams}
    batch["state"] = rng.standard_normal((5, 1, 7)).astype(np.float32)
    ab = agent.sample_actions(batch, seed=key)
    assert ab.shape == (5, 4) and rel_err(ab, O.sample_actions(ostate, ocfg, batch, seed=key).numpy()) < Q_TOL


def test_state_sac_update_high_utd_matches_oracle():
    """async_sac_state_sim (BASELINE config 1): state SAC, whole critic vmapped, 2000-step lr warm-up, UTD scan."""
    from oracle import drq as O
    from helpers import Box
    from serl_b200.utils.launcher import make_sac_agent
    import types
    S, A, B, utd = 10, 4, 32, 4
    rng = np.random.default_rng(0)
    agent = make_sac_agent(42, rng.standard_normal(S).astype(np.float32), rng.uniform(-1, 1, A).astype(np.float32))
    _perturb(agent, seed=4)
    agent._store.counts.fill_(700)            # inside the warm-up ramp so lr != 0
    ostate, ocfg = oracle_state_from_agent(agent), oracle_cfg_from_agent(agent)
    ocfg.discount = 0.99
    batch = dict(observations=rng.standard_normal((B, S)).astype(np.float32), next_observations=rng.standard_normal((B, S)).astype(np.float32),
                 actions=rng.uniform(-1, 1, (B, A)).astype(np.float32), rewards=rng.random(B).astype(np.float32),
                 masks=(rng.random(B) > 0.1).astype(np.float32), dones=np.zeros(B, bool))
    agent, info = agent.update_high_utd(batch, utd_ratio=utd)
    ob = dict(batch, observations={"state": batch["observations"]}, next_observations={"state": batch["next_observations"]})
    oinfo = O.update_high_utd(ostate, ocfg, ob, utd, augment=False)
    np.testing.assert_allclose(float(info["critic"]["critic_loss"]), oinfo["critic"]["critic_loss"], rtol=Q_TOL)
    np.testing.assert_allclose(float(info["actor"]["actor_loss"]), oinfo["actor"]["actor_loss"], rtol=Q_TOL, atol=1e-6)
    np.testing.assert_allclose(float(info["actor_lr"]), 3e-4 * (700 + utd) / 2000, rtol=1e-6)
    _compare_state(agent, ostate, oinfo, "state sac")
    assert agent.state.step == utd + 1
