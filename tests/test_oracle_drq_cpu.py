"""CPU: semantics of the oracle's restated SAC update (state agent, tiny) - the reference quirks the product must
reproduce: all three Adam transforms tick on every `update`, zero-gradient momentum
drift, polyak only after critic updates, key chain, float32 vs float64 agreement."""
import numpy as np
import torch

from oracle import drq as O
from oracle import jax_prng as P
from serl_b200.params import init_trainable, trainable_spec


def _setup(dtype=torch.float64, seed=0, S=5, A=3, E=4):
    rng = np.random.default_rng(seed)
    spec = trainable_spec((), S, A, E, pixel=False)
    params = {k: torch.as_tensor(v) for k, v in init_trainable(rng, spec, 1e-2).items()}
    for k in params:                                    # move biases / norms off their zero / one init
        params[k] = params[k] + 0.05 * torch.as_tensor(rng.standard_normal(params[k].shape).astype(np.float32))
    state = O.OracleState.create(params, P.prng_key(42), dtype)
    cfg = O.OracleConfig(cams=(), discount=0.99, target_entropy=-A / 2, ensemble=E, subsample=2, pixel=False)
    B = 16
    batch = dict(observations={"state": rng.standard_normal((B, S)).astype(np.float32)},
                 next_observations={"state": rng.standard_normal((B, S)).astype(np.float32)},
                 actions=rng.uniform(-1, 1, (B, A)).astype(np.float32), rewards=rng.random(B).astype(np.float32),
                 masks=(rng.random(B) > 0.2).astype(np.float32))
    return state, cfg, batch, A


def _step(state, cfg, batch, A, nets):
    rnd, new_rng = O.derive_update_randomness(state.rng, 16, A, (), False, cfg.ensemble, cfg.subsample, nets=tuple(nets))
    return O.update(state, cfg, batch, rnd, frozenset(nets), state.params["modules_temperature/lagrange"].dtype, new_rng)


def test_all_three_optimizers_tick_and_zero_grad_momentum_drifts():
    state, cfg, batch, A = _setup()
    actor_keys = [k for k in state.params if k.startswith("modules_actor")]
    p0 = {k: v.clone() for k, v in state.params.items()}
    _step(state, cfg, batch, A, {"critic"})
    assert [state.opt[n]["count"] for n in ("actor", "critic", "temperature")] == [1, 1, 1]
    for k in actor_keys:                                 # no actor gradient yet: moments are 0 -> exactly no movement
        assert torch.equal(state.params[k], p0[k])
    assert not torch.equal(state.params["modules_critic/Dense_0/kernel"], p0["modules_critic/Dense_0/kernel"])
    _step(state, cfg, batch, A, {"actor", "temperature"})
    p1 = {k: v.clone() for k, v in state.params.items()}
    t1 = {k: v.clone() for k, v in state.target_params.items()}
    _step(state, cfg, batch, A, {"critic"})              # critic-only step: actor + temperature still move (momentum)
    moved = sum(float((state.params[k] - p1[k]).abs().max()) for k in actor_keys)
    assert moved > 0
    assert float((state.params["modules_temperature/lagrange"] - p1["modules_temperature/lagrange"]).abs()) > 0
    assert [state.opt[n]["count"] for n in ("actor", "critic", "temperature")] == [3, 3, 3]
    # polyak over the whole tree after the critic update, tau = 0.005
    k = "modules_critic/network/Dense_0/kernel"
    torch.testing.assert_close(state.target_params[k], state.params[k] * 0.005 + t1[k] * 0.995)


def test_polyak_skipped_without_critic_and_rng_chain():
    state, cfg, batch, A = _setup()
    t0 = {k: v.clone() for k, v in state.target_params.items()}
    r0 = state.rng.copy()
    _step(state, cfg, batch, A, {"actor", "temperature"})
    for k in t0:
        assert torch.equal(state.target_params[k], t0[k])            # sac.py:284: target only moves with the critic
    np.testing.assert_array_equal(state.rng, P.split(r0)[0])         # sac.py:288


def test_float32_oracle_tracks_float64():
    s64, cfg, batch, A = _setup(torch.float64)
    s32, _, _, _ = _setup(torch.float32)
    i64 = _step(s64, cfg, batch, A, {"critic"})
    i32 = _step(s32, cfg, batch, A, {"critic"})
    assert abs(i64["critic"]["critic_loss"] - i32["critic"]["critic_loss"]) < 1e-5 * max(1.0, abs(i64["critic"]["critic_loss"]))
    q64, q32 = i64["critic"]["_q"], i32["critic"]["_q"].double()
    assert float((q64 - q32).abs().max()) < 1e-5 * float(q64.abs().max())


def test_subsample_is_with_replacement_and_min_over_two():
    state, cfg, batch, A = _setup()
    rnd, new_rng = O.derive_update_randomness(state.rng, 16, A, (), False, cfg.ensemble, cfg.subsample, nets=("critic",))
    rnd.critic.subsample = np.array([1, 1], np.int32)                # a repeated member is legal (randint with replacement)
    info = O.update(state, cfg, batch, rnd, frozenset({"critic"}), torch.float64, new_rng)
    assert np.isfinite(info["critic"]["critic_loss"])
    assert info["critic"]["_target_q"].shape == (16,)


def test_actor_loss_differentiates_the_proprio_encoder_but_not_the_image_heads():
    """Audit of the reference's stop_gradient sites for the pixel agent:
      * Policy.__call__ calls encoder(obs, train, stop_gradient=True)      (networks/actor_critic_nets.py:185)
      * EncodingWrapper stops the gradient at each per-camera image embedding (common/encoding.py:48-49) and NOT at the
        proprio Dense -> LayerNorm -> tanh that follows                    (common/encoding.py:55-70)
      * policy_loss_fn differentiates w.r.t. the full tree                 (agents/continuous/sac.py:198-200)
    so d(actor loss) reaches modules_actor/encoder/{Dense_0,LayerNorm_0} and is exactly zero for encoder_<cam>/*.
    The autograd result is checked against central finite differences of the literal formulation (perturbing ONLY the
    `grad_params` copy, like jax.grad does: forward_critic inside the actor loss reads self.state.params)."""
    import torch.nn.functional as F
    from serl_b200.params import ENC, init_trainable, trainable_spec
    rng = np.random.default_rng(3)
    cams, S, A, E, B = ("front",), 7, 4, 3, 5
    spec = trainable_spec(cams, S, A, E, pixel=True)
    params = {k: torch.as_tensor(v).double() for k, v in init_trainable(rng, spec, 1e-2).items()}
    for k in params:
        params[k] = params[k] + 0.05 * torch.as_tensor(rng.standard_normal(params[k].shape))
    cfg = O.OracleConfig(cams=cams, ensemble=E, subsample=2, pixel=True)
    feats = {"front": torch.as_tensor(np.abs(rng.standard_normal((B, 4, 4, 512))))}     # frozen-trunk output (stop_gradient)
    state_obs = rng.standard_normal((B, 1, S))
    eps = torch.as_tensor(rng.standard_normal((B, A)))
    drop = {"front": torch.as_tensor(rng.random((B, 4096)) < 0.9)}
    lam = "modules_temperature/lagrange"

    def actor_loss(grad_params):
        enc = O.encode(grad_params, cams, feats, torch.as_tensor(state_obs), drop, stop_gradient=True)
        mu, sd = O.policy_forward(grad_params, enc)
        a, logp = O.tanh_normal_sample_logp(mu, sd, eps)
        with torch.no_grad():
            enc_c = O.encode(params, cams, feats, torch.as_tensor(state_obs), None)
        q = O.critic_forward(params, enc_c, a, True).mean(dim=0)
        return -(q - F.softplus(params[lam]).detach() * logp).mean()

    leaves = {k: v.clone().requires_grad_(True) for k, v in params.items()}
    grads = dict(zip(leaves, torch.autograd.grad(actor_loss(leaves), list(leaves.values()), allow_unused=True)))
    for k, g in grads.items():
        if "/encoder_" in k or k.startswith("modules_critic") or k == lam:
            assert g is None or float(g.abs().max()) == 0.0, k
    for k in (f"{ENC}/Dense_0/kernel", f"{ENC}/Dense_0/bias", f"{ENC}/LayerNorm_0/scale", f"{ENC}/LayerNorm_0/bias"):
        g = grads[k]
        assert g is not None and float(g.abs().max()) > 0, k
        idx = tuple(int(i) for i in np.unravel_index(int(g.abs().argmax()), g.shape))
        h = 1e-6
        vals = []
        for sgn in (+1, -1):
            pert = {kk: v.clone() for kk, v in params.items()}
            pert[k][idx] += sgn * h
            vals.append(float(actor_loss(pert)))
        fd = (vals[0] - vals[1]) / (2 * h)
        assert abs(fd - float(g[idx])) < 1e-6 * max(1.0, abs(fd)), (k, fd, float(g[idx]))
    # and the oracle's update() routes it to the ACTOR transform: Adam moments of the actor tx become non-zero there
    st = O.OracleState.create({k: v.clone() for k, v in params.items()}, P.prng_key(1), torch.float64)
    batch = dict(observations={"state": state_obs, "front": None}, next_observations={"state": state_obs, "front": None},
                 rewards=np.zeros(B, np.float32), masks=np.ones(B, np.float32), actions=np.zeros((B, A), np.float32))
    rnd = O.UpdateRandomness(actor=O.LossRandomness(eps=eps.numpy(), dropout={"front": drop["front"].numpy()}))
    real_features = O._features
    O._features = lambda *a, **k: feats
    try:
        O.update(st, cfg, batch, rnd, frozenset({"actor"}), torch.float64)
    finally:
        O._features = real_features
    assert float(st.opt["actor"]["mu"][f"{ENC}/Dense_0/kernel"].abs().max()) > 0
    assert float(st.opt["critic"]["mu"][f"{ENC}/Dense_0/kernel"].abs().max()) == 0
    assert float(st.opt["actor"]["mu"][f"{ENC}/encoder_front/Dense_0/kernel"].abs().max()) == 0
