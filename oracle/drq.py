"""ORACLE (test infrastructure, not product): PyTorch-CPU restatement of the reference's
DrQ / SAC gradient step, float64 ("truth") or float32.

Follows (file:line relative to /root/reference/serl_launcher/serl_launcher):
  vision/resnet_v1.py:189-286   frozen ResNet-10 trunk (normalise, conv7x7/2 pad 3, GN(4), ReLU,
                                max_pool 3x3/2 SAME, 4 basic blocks, stop_gradient)
  vision/resnet_v1.py:129-156   ResNetBlock (conv-GN-ReLU-conv-GN, 1x1 proj+GN, ReLU(res+y))
  vision/resnet_v1.py:81-116    SpatialLearnedEmbeddings
  vision/resnet_v1.py:324-376   PreTrainedResNetEncoder head (SLE, Dropout(0.1), Dense, LayerNorm, tanh)
  common/encoding.py:26-72      EncodingWrapper (per-camera encode, concat, proprio Dense/LN/tanh)
  networks/mlp.py:10-32         MLP (Dense -> LayerNorm -> tanh per layer, activate_final)
  networks/actor_critic_nets.py:49-73,156-164   Critic (+ vmapped ensemble backbone, shared value head)
  networks/actor_critic_nets.py:167-272         Policy, TanhMultivariateNormalDiag
  networks/lagrange.py:9-78     GeqLagrangeMultiplier (softplus parameterisation)
  agents/continuous/sac.py:118-299,301-320,544-596   losses, update, sample_actions, update_high_utd
  agents/continuous/drq.py:244-328                   augmentation + update_critics / update_high_utd
  common/common.py:124-221      target_update, apply_gradients (3 full-tree Adam txs summed), apply_loss_fns
  common/optimizers.py:6-56     Adam + warmup schedule (optax semantics restated in `adam_tx_update`)
Third-party arithmetic restated from published definitions (not under /root/reference):
flax.linen Dense/LayerNorm(eps 1e-6, fast variance)/GroupNorm/Conv(SAME)/max_pool/Dropout,
optax.adam (b1 .9, b2 .999, eps 1e-8, bias-corrected), distrax tanh-Gaussian log-prob.

PARITY PIN STATUS: **parity unpinned** for the network arithmetic - the reference ships no
golden vectors and jax/flax/optax/distrax cannot be installed here.  Only the PRNG
(oracle/jax_prng.py) and the replay layout (oracle/replay.py) are pinned to external vectors.

Parameters are a flat dict {"modules_actor/encoder/...": tensor} in the Flax tree layout
(SURVEY.md Appendix D).  All randomness is explicit (`StepRandomness`) and can be derived from a
JAX-style key with `derive_*` below (key chain of SURVEY.md Appendix A.2/A.3).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline/reference legs may import this.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, Optional, Sequence

import numpy as np
import torch
import torch.nn.functional as F

from . import jax_prng as P
from .replay import random_shift

ENC = "modules_actor/encoder"
IMAGENET_MEAN = (0.485, 0.456, 0.406)
IMAGENET_STD = (0.229, 0.224, 0.225)
STAGES = ((64, 1), (128, 2), (256, 2), (512, 2))       # (filters, stride) of ResNet-10's 4 blocks


# ------------------------------------------------------------------------------------------
# layers
# ------------------------------------------------------------------------------------------
def _norm_fast_var(x, dims, eps):
    """flax normalisation statistics: var = E[x^2] - E[x]^2 (use_fast_variance=True), clipped at 0."""
    mean = x.mean(dim=dims, keepdim=True)
    var = ((x * x).mean(dim=dims, keepdim=True) - mean * mean).clamp_min(0)
    return (x - mean) * torch.rsqrt(var + eps)


def layer_norm(x, scale, bias, eps=1e-6):
    return _norm_fast_var(x, (-1,), eps) * scale + bias


def group_norm_nhwc(x, scale, bias, groups=4, eps=1e-5):
    n, h, w, c = x.shape
    xg = x.reshape(n, h, w, groups, c // groups)
    xg = _norm_fast_var(xg, (1, 2, 4), eps)
    return xg.reshape(n, h, w, c) * scale + bias


def conv_nhwc(x, kernel_hwio, stride, pad_lo, pad_hi):
    """NHWC conv with explicit asymmetric zero padding (XLA SAME on even sizes pads low 0 / high 1)."""
    xin = x.permute(0, 3, 1, 2)
    xin = F.pad(xin, (pad_lo, pad_hi, pad_lo, pad_hi))
    w = kernel_hwio.permute(3, 2, 0, 1)
    return F.conv2d(xin, w, stride=stride).permute(0, 2, 3, 1)


def same_pads(size, k, stride):
    out = -(-size // stride)
    total = max((out - 1) * stride + k - size, 0)
    return total // 2, total - total // 2


def max_pool_3x3_s2_same(x):
    n, h, w, c = x.shape
    lo, hi = same_pads(h, 3, 2)
    xin = F.pad(x.permute(0, 3, 1, 2), (lo, hi, lo, hi), value=float("-inf"))
    return F.max_pool2d(xin, 3, 2).permute(0, 2, 3, 1)


def trunk_forward(params: Dict[str, torch.Tensor], cam: str, images_u8: torch.Tensor, dtype) -> torch.Tensor:
    """(N,128,128,3*T) uint8 -> (N,4,4,512).  resnet_v1.py:217-286; stop_gradient -> no_grad."""
    pre = f"{ENC}/encoder_{cam}/pretrained_encoder"
    g = lambda k: params[f"{pre}/{k}"].to(dtype)
    with torch.no_grad():
        c_in = images_u8.shape[-1]
        reps = c_in // 3
        mean = torch.tensor(IMAGENET_MEAN * reps, dtype=dtype)
        std = torch.tensor(IMAGENET_STD * reps, dtype=dtype)
        x = (images_u8.to(dtype) / 255.0 - mean) / std
        x = conv_nhwc(x, g("conv_init/kernel"), 2, 3, 3)
        x = group_norm_nhwc(x, g("norm_init/scale"), g("norm_init/bias")).relu()
        x = max_pool_3x3_s2_same(x)
        for i, (filters, stride) in enumerate(STAGES):
            b = f"ResNetBlock_{i}"
            res = x
            lo, hi = same_pads(x.shape[1], 3, stride)
            y = conv_nhwc(x, g(f"{b}/Conv_0/kernel"), stride, lo, hi)
            y = group_norm_nhwc(y, g(f"{b}/MyGroupNorm_0/scale"), g(f"{b}/MyGroupNorm_0/bias")).relu()
            y = conv_nhwc(y, g(f"{b}/Conv_1/kernel"), 1, 1, 1)
            y = group_norm_nhwc(y, g(f"{b}/MyGroupNorm_1/scale"), g(f"{b}/MyGroupNorm_1/bias"))
            if res.shape != y.shape:
                res = conv_nhwc(res, g(f"{b}/conv_proj/kernel"), stride, 0, 0)
                res = group_norm_nhwc(res, g(f"{b}/norm_proj/scale"), g(f"{b}/norm_proj/bias"))
            x = (res + y).relu()
    return x


def encode(params, cams: Sequence[str], feats: Dict[str, torch.Tensor], state: torch.Tensor,
           dropout_masks: Optional[Dict[str, torch.Tensor]] = None, stop_gradient: bool = False) -> torch.Tensor:
    """encoding.py:26-72 with the per-camera head of resnet_v1.py:340-374.
    feats[cam] (B,4,4,512); state (B,T,S); dropout_masks[cam] (B,4096) bool keep-mask or None.
    stop_gradient (encoding.py:48-49): applied to each per-camera IMAGE embedding only - the proprio
    Dense -> LayerNorm -> tanh below (:55-70) stays differentiable."""
    outs = []
    for cam in cams:
        pre = f"{ENC}/encoder_{cam}"
        k = params[f"{pre}/SpatialLearnedEmbeddings_0/kernel"]             # (4,4,512,8)
        f = feats[cam].to(k.dtype)
        sle = torch.einsum("bhwc,hwcf->bcf", f, k).reshape(f.shape[0], -1)  # index c*8+f
        if dropout_masks is not None:
            keep = 0.9
            sle = torch.where(dropout_masks[cam], sle / keep, torch.zeros_like(sle))
        z = sle @ params[f"{pre}/Dense_0/kernel"] + params[f"{pre}/Dense_0/bias"]
        z = layer_norm(z, params[f"{pre}/LayerNorm_0/scale"], params[f"{pre}/LayerNorm_0/bias"])
        img = torch.tanh(z)
        outs.append(img.detach() if stop_gradient else img)               # encoding.py:48-49
    s = state.reshape(state.shape[0], -1).to(outs[0].dtype)
    z = s @ params[f"{ENC}/Dense_0/kernel"] + params[f"{ENC}/Dense_0/bias"]
    z = layer_norm(z, params[f"{ENC}/LayerNorm_0/scale"], params[f"{ENC}/LayerNorm_0/bias"])
    outs.append(torch.tanh(z))
    return torch.cat(outs, dim=-1)


def mlp2(params, prefix, x, ensemble: bool):
    """mlp.py:22-31, hidden [256,256], LayerNorm, tanh, activate_final.  Ensemble params have a leading E axis."""
    for i in range(2):
        w, b = params[f"{prefix}/Dense_{i}/kernel"], params[f"{prefix}/Dense_{i}/bias"]
        sc, bi = params[f"{prefix}/LayerNorm_{i}/scale"], params[f"{prefix}/LayerNorm_{i}/bias"]
        if ensemble:
            x = (torch.einsum("bi,eio->ebo", x, w) if x.dim() == 2 else torch.einsum("ebi,eio->ebo", x, w)) + b[:, None, :]
            x = torch.tanh(layer_norm(x, sc[:, None, :], bi[:, None, :]))
        else:
            x = torch.tanh(layer_norm(x @ w + b, sc, bi))
    return x


def critic_forward(params, enc, actions, pixel_agent: bool = True):
    """actor_critic_nets.py:57-73 -> (E,B).  Pixel agent: ensembled backbone + ONE shared value head
    (drq.py:201-207); state agent: whole critic vmapped incl. head (sac.py:523-524)."""
    x = torch.cat([enc, actions.to(enc.dtype)], dim=-1)
    h = mlp2(params, "modules_critic/network", x, ensemble=True)            # (E,B,256)
    w, b = params["modules_critic/Dense_0/kernel"], params["modules_critic/Dense_0/bias"]
    if pixel_agent:
        return (h @ w + b).squeeze(-1)
    return (torch.einsum("ebi,eio->ebo", h, w) + b[:, None, :]).squeeze(-1)


def policy_forward(params, enc, std_min=1e-5, std_max=5.0):
    """actor_critic_nets.py:178-227, std_parameterization="exp" -> (means, stds)."""
    h = mlp2(params, "modules_actor/network", enc, ensemble=False)
    means = h @ params["modules_actor/Dense_0/kernel"] + params["modules_actor/Dense_0/bias"]
    log_stds = h @ params["modules_actor/Dense_1/kernel"] + params["modules_actor/Dense_1/bias"]
    return means, torch.clamp(torch.exp(log_stds), std_min, std_max)


def tanh_normal_sample_logp(means, stds, eps):
    """distrax Transformed(MultivariateNormalDiag, Block(Tanh)).sample_and_log_prob."""
    eps = eps.to(means.dtype)
    u = means + stds * eps
    a = torch.tanh(u)
    z = (u - means) / stds
    base = (-0.5 * z * z - torch.log(stds) - 0.5 * math.log(2 * math.pi)).sum(-1)
    fldj = (2.0 * (math.log(2.0) - u - F.softplus(-2.0 * u))).sum(-1)
    return a, base - fldj


# ------------------------------------------------------------------------------------------
# randomness (explicit) and its derivation from a JAX key
# ------------------------------------------------------------------------------------------
@dataclass
class LossRandomness:
    eps: np.ndarray                                   # (B,A) float32
    dropout: Dict[str, np.ndarray]                    # cam -> (B,4096) bool keep-mask
    subsample: Optional[np.ndarray] = None            # (2,) int32, critic only


@dataclass
class UpdateRandomness:
    critic: Optional[LossRandomness] = None
    actor: Optional[LossRandomness] = None
    temperature: Optional[LossRandomness] = None


def _dropout_masks(key, cams, B):
    """Repo spec (flax's make_rng path-hash folding is version-coupled and not reproducible here):
    camera j's keep-mask = bernoulli(fold_in(key, j), 0.9, (B, 4096))."""
    return {cam: P.bernoulli(P.fold_in(key, j), 0.9, (B, 4096)) for j, cam in enumerate(cams)}


def derive_update_randomness(rng, B, A, cams, pixel: bool, ensemble=10, subsample=2,
                             nets=("critic", "actor", "temperature")):
    """Key chain of SACAgent.update (sac.py:243-299 -> common.py:198-200 -> loss fns).
    Returns (UpdateRandomness, new_state_rng)."""
    _, k_actor, k_critic, k_temp = P.split(rng, 4)          # sorted-key order actor, critic, temperature
    out = UpdateRandomness()
    if "critic" in nets:
        c1, k_na = P.split(k_critic)                        # sac.py:137
        c2, k_sub = P.split(c1)                             # sac.py:152
        out.critic = LossRandomness(eps=P.normal(k_na, (B, A)),
                                    dropout=_dropout_masks(k_na, cams, B) if pixel else {},
                                    subsample=P.randint(k_sub, (subsample,), 0, ensemble))
    if "actor" in nets:
        _, k_p, k_s, _k_c = P.split(k_actor, 4)             # sac.py:197
        out.actor = LossRandomness(eps=P.normal(k_s, (B, A)),
                                   dropout=_dropout_masks(k_p, cams, B) if pixel else {})
    if "temperature" in nets:
        _, k = P.split(k_temp)                              # sac.py:224
        out.temperature = LossRandomness(eps=P.normal(k, (B, A)),
                                         dropout=_dropout_masks(k, cams, B) if pixel else {})
    return out, P.split(rng)[0]                             # sac.py:288


def derive_augmentation(rng, n_frames):
    """drq.py:307-310: rng, obs_rng, next_rng = split(state.rng, 3).  Returns (new_rng, off_obs, off_next)."""
    r1, k_obs, k_next = P.split(rng, 3)
    return r1, P.crop_offsets(k_obs, n_frames), P.crop_offsets(k_next, n_frames)


# ------------------------------------------------------------------------------------------
# optimiser: optax.inject_hyperparams(adam) restated; full-tree, one state per tx
# ------------------------------------------------------------------------------------------
def lr_schedule(count: int, lr: float, warmup: int) -> float:
    """optimizers.py:23-29: join_schedules([linear(0 -> lr, warmup), constant(lr)], [warmup])."""
    if count < warmup:
        return lr * (count / warmup)
    return lr


def adam_tx_update(grads, opt, lr, b1=0.9, b2=0.999, eps=1e-8):
    """One tx.update over the whole tree.  opt = {"count": int, "mu": {...}, "nu": {...}}."""
    t = opt["count"] + 1
    updates = {}
    for k, g in grads.items():
        mu = b1 * opt["mu"][k] + (1 - b1) * g
        nu = b2 * opt["nu"][k] + (1 - b2) * g * g
        opt["mu"][k], opt["nu"][k] = mu, nu
        mu_hat = mu / (1 - b1 ** t)
        nu_hat = nu / (1 - b2 ** t)
        updates[k] = -lr * mu_hat / (torch.sqrt(nu_hat) + eps)
    opt["count"] = t
    return updates


# ------------------------------------------------------------------------------------------
# the agent state + update
# ------------------------------------------------------------------------------------------
@dataclass
class OracleState:
    params: Dict[str, torch.Tensor]
    target_params: Dict[str, torch.Tensor]
    opt: Dict[str, dict]
    rng: np.ndarray
    step: int = 0

    @classmethod
    def create(cls, params, rng, dtype):
        p = {k: v.detach().to(dtype).clone() for k, v in params.items()}
        zeros = lambda: {k: torch.zeros_like(v) for k, v in p.items()}
        return cls(params=p, target_params={k: v.clone() for k, v in p.items()},
                   opt={n: {"count": 0, "mu": zeros(), "nu": zeros()} for n in ("actor", "critic", "temperature")},
                   rng=np.asarray(rng, dtype=np.uint32).copy())


@dataclass
class OracleConfig:
    cams: Sequence[str] = ()
    discount: float = 0.96
    tau: float = 0.005
    target_entropy: float = -2.0
    ensemble: int = 10
    subsample: Optional[int] = 2
    backup_entropy: bool = False
    lr: float = 3e-4
    warmup: Dict[str, int] = field(default_factory=lambda: {"actor": 0, "critic": 0, "temperature": 0})
    pixel: bool = True


def _features(state: OracleState, cfg: OracleConfig, obs: dict, dtype):
    """Frozen-trunk features per camera.  "B T H W C -> B H W (T C)" (encoding.py:41-44)."""
    feats = {}
    for cam in cfg.cams:
        img = torch.as_tensor(np.asarray(obs[cam]))
        b, t, h, w, c = img.shape
        img = img.permute(0, 2, 3, 1, 4).reshape(b, h, w, t * c)
        feats[cam] = trunk_forward(state.params, cam, img, dtype)
    return feats


def _enc(params, cfg, feats, obs_state, masks, stop_gradient=False):
    if not cfg.pixel:
        return torch.as_tensor(np.asarray(obs_state)).to(next(iter(params.values())).dtype)
    m = None if masks is None else {c: torch.as_tensor(v) for c, v in masks.items()}
    return encode(params, cfg.cams, feats, torch.as_tensor(np.asarray(obs_state)), m, stop_gradient)


def update(state: OracleState, cfg: OracleConfig, batch: dict, rnd: UpdateRandomness,
           nets=frozenset({"actor", "critic", "temperature"}), dtype=torch.float64, new_rng=None):
    """SACAgent.update (sac.py:243-299) on an already-augmented, unpacked batch.  Mutates `state`."""
    p0 = state.params
    leaves = {k: v.clone().requires_grad_(True) for k, v in p0.items()}
    const = p0
    obs, nobs = batch["observations"], batch["next_observations"]
    rewards = torch.as_tensor(np.asarray(batch["rewards"])).to(dtype)
    masks = torch.as_tensor(np.asarray(batch["masks"])).to(dtype)
    actions = torch.as_tensor(np.asarray(batch["actions"])).to(dtype)
    feats_o = _features(state, cfg, obs, dtype) if cfg.pixel else None
    feats_n = _features(state, cfg, nobs, dtype) if cfg.pixel else None
    info, grads = {}, {}
    zero_grads = lambda: {k: torch.zeros_like(v) for k, v in p0.items()}

    def grad_of(loss):
        gs = torch.autograd.grad(loss, list(leaves.values()), allow_unused=True)
        return {k: (torch.zeros_like(p0[k]) if g is None else g) for k, g in zip(leaves, gs)}

    lam = "modules_temperature/lagrange"
    if "critic" in nets:                                              # sac.py:134-191
        r = rnd.critic
        with torch.no_grad():
            enc_n = _enc(const, cfg, feats_n, nobs["state"], r.dropout if cfg.pixel else None)
            mu, sd = policy_forward(const, enc_n)
            a_n, logp_n = tanh_normal_sample_logp(mu, sd, torch.as_tensor(r.eps))
            enc_nt = _enc(state.target_params, cfg, feats_n, nobs["state"], None)
            q_next = critic_forward(state.target_params, enc_nt, a_n, cfg.pixel)       # (E,B)
            if cfg.subsample is not None:
                q_next = q_next[torch.as_tensor(np.asarray(r.subsample), dtype=torch.long)]
            target_q = rewards + cfg.discount * masks * q_next.min(dim=0).values
            if cfg.backup_entropy:
                target_q = target_q - F.softplus(const[lam]) * logp_n
        enc_o = _enc(leaves, cfg, feats_o, obs["state"], None)
        q = critic_forward(leaves, enc_o, actions, cfg.pixel)
        loss = ((q - target_q[None]) ** 2).mean()
        grads["critic"] = grad_of(loss)
        info["critic"] = {"critic_loss": loss.item(), "predicted_qs": q.mean().item(),
                          "target_qs": target_q.mean().item(), "_q": q.detach(), "_target_q": target_q}
    else:
        grads["critic"] = zero_grads()
    if "actor" in nets:                                               # sac.py:193-221
        r = rnd.actor
        temperature = F.softplus(const[lam]).detach()
        # forward_policy(obs, grad_params=params) (sac.py:198-200) -> Policy.__call__: encoder(obs, train, stop_gradient=True)
        # (actor_critic_nets.py:185).  EncodingWrapper stops the gradient at the per-camera image embeddings ONLY
        # (encoding.py:48-49); the proprio Dense/LayerNorm (:55-70) is differentiated w.r.t. `leaves`.
        enc_o = _enc(leaves, cfg, feats_o, obs["state"], r.dropout if cfg.pixel else None, stop_gradient=True)
        mu, sd = policy_forward(leaves, enc_o)
        a, logp = tanh_normal_sample_logp(mu, sd, torch.as_tensor(r.eps))
        with torch.no_grad():
            enc_c = _enc(const, cfg, feats_o, obs["state"], None)
        qa = critic_forward(const, enc_c, a, cfg.pixel).mean(dim=0)
        loss = -(qa - temperature * logp).mean()
        grads["actor"] = grad_of(loss)
        info["actor"] = {"actor_loss": loss.item(), "temperature": temperature.item(),
                         "entropy": (-logp.mean()).item(), "_actions": a.detach(), "_log_probs": logp.detach()}
    else:
        grads["actor"] = zero_grads()
    if "temperature" in nets:                                         # sac.py:223-234
        r = rnd.temperature
        with torch.no_grad():
            enc_n = _enc(const, cfg, feats_n, nobs["state"], r.dropout if cfg.pixel else None)
            mu, sd = policy_forward(const, enc_n)
            _, logp_n = tanh_normal_sample_logp(mu, sd, torch.as_tensor(r.eps))
            entropy = -logp_n.mean()
        loss = F.softplus(leaves[lam]) * (entropy - cfg.target_entropy)
        grads["temperature"] = grad_of(loss)
        info["temperature"] = {"temperature_loss": loss.item()}
    else:
        grads["temperature"] = zero_grads()

    # common.py:136-168: every tx runs over the whole tree; updates are summed
    total = {k: torch.zeros_like(v) for k, v in p0.items()}
    for name in ("actor", "critic", "temperature"):
        lr = lr_schedule(state.opt[name]["count"], cfg.lr, cfg.warmup[name])
        upd = adam_tx_update(grads[name], state.opt[name], lr)
        info[f"{name}_lr"] = lr
        for k in total:
            total[k] = total[k] + upd[k]
    state.params = {k: (p0[k] + total[k]).detach() for k in p0}
    state.step += 1
    if "critic" in nets:                                              # common.py:124-134
        state.target_params = {k: state.params[k] * cfg.tau + state.target_params[k] * (1 - cfg.tau)
                               for k in state.params}
    if new_rng is not None:
        state.rng = np.asarray(new_rng, dtype=np.uint32)
    info["_grads"] = grads
    return info


def _augment(batch: dict, cams, off_obs, off_next):
    """drq.py:244-253 on an unpacked batch: the SAME offsets for every camera of a sample."""
    out = dict(batch)
    obs, nobs = dict(batch["observations"]), dict(batch["next_observations"])
    for cam in cams:
        for d, off in ((obs, off_obs), (nobs, off_next)):
            x = np.asarray(d[cam])
            b, t = x.shape[:2]
            d[cam] = random_shift(x.reshape(b * t, *x.shape[2:]), off).reshape(x.shape)
    out["observations"], out["next_observations"] = obs, nobs
    return out


def update_critics(state: OracleState, cfg: OracleConfig, batch_unpacked: dict, dtype=torch.float64):
    """DrQAgent.update_critics (drq.py:296-328), seeded from state.rng."""
    B = np.asarray(batch_unpacked["rewards"]).shape[0]
    A = np.asarray(batch_unpacked["actions"]).shape[-1]
    T = np.asarray(batch_unpacked["observations"][cfg.cams[0]]).shape[1]
    r1, off_o, off_n = derive_augmentation(state.rng, B * T)
    aug = _augment(batch_unpacked, cfg.cams, off_o, off_n)
    rnd, new_rng = derive_update_randomness(r1, B, A, cfg.cams, cfg.pixel, cfg.ensemble, cfg.subsample or 0,
                                            nets=("critic",))
    info = update(state, cfg, aug, rnd, frozenset({"critic"}), dtype, new_rng)
    info["_aug"] = aug
    return info


def update_high_utd(state: OracleState, cfg: OracleConfig, batch_unpacked: dict, utd_ratio: int,
                    dtype=torch.float64, augment: bool = True):
    """DrQAgent.update_high_utd (drq.py:255-294) -> SACAgent.update_high_utd (sac.py:544-596)."""
    B = np.asarray(batch_unpacked["rewards"]).shape[0]
    A = np.asarray(batch_unpacked["actions"]).shape[-1]
    assert B % utd_ratio == 0
    mb = B // utd_ratio
    rng = state.rng
    batch = batch_unpacked
    if augment and cfg.pixel:
        T = np.asarray(batch["observations"][cfg.cams[0]]).shape[1]
        rng, off_o, off_n = derive_augmentation(rng, B * T)
        batch = _augment(batch, cfg.cams, off_o, off_n)

    def rows(x, lo, hi):
        return {k: rows(v, lo, hi) for k, v in x.items()} if isinstance(x, dict) else np.asarray(x)[lo:hi]

    crit = []
    for i in range(utd_ratio):
        rnd, new_rng = derive_update_randomness(rng, mb, A, cfg.cams, cfg.pixel, cfg.ensemble, cfg.subsample or 0,
                                                nets=("critic",))
        crit.append(update(state, cfg, rows(batch, i * mb, (i + 1) * mb), rnd, frozenset({"critic"}), dtype, new_rng))
        rng = new_rng
    rnd, new_rng = derive_update_randomness(rng, B, A, cfg.cams, cfg.pixel, cfg.ensemble, cfg.subsample or 0,
                                            nets=("actor", "temperature"))
    at = update(state, cfg, batch, rnd, frozenset({"actor", "temperature"}), dtype, new_rng)
    info = {"critic": {k: float(np.mean([c["critic"][k] for c in crit]))
                       for k in ("critic_loss", "predicted_qs", "target_qs")},
            "actor": at["actor"], "temperature": at["temperature"]}
    for n in ("actor", "critic", "temperature"):
        info[f"{n}_lr"] = at[f"{n}_lr"]          # sac.py:592: {**critic_infos, **actor_temp_infos}
    info["_aug"] = batch
    merged = {n: {k: sum(c["_grads"][n][k].abs() for c in crit) + at["_grads"][n][k].abs() for k in at["_grads"][n]}
              for n in at["_grads"]}
    info["_grads"] = at["_grads"]
    info["_grads_abs_all_calls"] = merged
    return info


def sample_actions(state: OracleState, cfg: OracleConfig, obs: dict, seed=None, argmax=False, dtype=torch.float64):
    """SACAgent.sample_actions (sac.py:301-320): train=False -> no dropout."""
    with torch.no_grad():
        feats = _features(state, cfg, obs, dtype) if cfg.pixel else None
        enc = _enc(state.params, cfg, feats, obs["state"], None)
        mu, sd = policy_forward(state.params, enc)
        if argmax:
            return torch.tanh(mu)
        eps = P.normal(np.asarray(seed, dtype=np.uint32), tuple(mu.shape))
        return torch.tanh(mu + sd * torch.as_tensor(eps).to(mu.dtype))
