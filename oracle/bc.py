"""CPU restatement of the reference's BCAgent.update (test infrastructure only - never imported by serl_b200/).

Follows agents/continuous/bc.py:36-76 with the networks `make_bc_agent` builds (utils/launcher.py:26-47): Policy
(networks/actor_critic_nets.py:167-227: encoder(..., stop_gradient=True), MLP [256, 256] with tanh and no LayerNorm
(networks/mlp.py:10-32), exp std clipped to [std_min, std_max], MultivariateNormalDiag), one optax.adam(3e-4)
(bc.py:199).  Encoder / trunk / Adam algebra are the functions of oracle/drq.py.  PARITY UNPINNED like oracle/drq.py
(jax / flax / distrax are not installable here): the layer definitions are restated from their published forms.
"""
from __future__ import annotations

import math

import numpy as np
import torch

from . import drq as O
from .jax_prng import split

ENC = O.ENC


def bc_forward(params, cams, feats, state, masks, std_min=1e-5, std_max=5.0):
    enc = O.encode(params, cams, feats, state, masks, stop_gradient=True)      # actor_critic_nets.py:185 + encoding.py:48-49
    n = "modules_actor/network"
    h = torch.tanh(enc @ params[f"{n}/Dense_0/kernel"] + params[f"{n}/Dense_0/bias"])
    h = torch.tanh(h @ params[f"{n}/Dense_1/kernel"] + params[f"{n}/Dense_1/bias"])
    mu = h @ params["modules_actor/Dense_0/kernel"] + params["modules_actor/Dense_0/bias"]
    ls = h @ params["modules_actor/Dense_1/kernel"] + params["modules_actor/Dense_1/bias"]
    return mu, torch.clamp(torch.exp(ls), std_min, std_max)


def update(params, opt, rng, cams, batch, dropout_masks=None, lr=3e-4, std_min=1e-5, std_max=5.0, dtype=torch.float64):
    """One BCAgent.update.  params: flat {path: tensor} incl. the frozen trunk; opt = {"count", "mu", "nu"} over the trainable
    leaves; returns (new_params, opt, new_rng, info, grads).  dropout_masks None -> keyed masks (repo spec, oracle/drq.py)."""
    new_rng, k = split(np.asarray(rng, np.uint32), 2)
    drop_key = split(k, 2)[1]
    obs = batch["observations"]
    actions = torch.as_tensor(np.asarray(batch["actions"])).to(dtype)
    B = actions.shape[0]
    p = {kk: v.detach().to(dtype) for kk, v in params.items()}
    train = {kk: v.clone().requires_grad_(True) for kk, v in p.items() if "pretrained_encoder" not in kk}
    full = {**p, **train}
    feats = {}
    for cam in cams:
        img = torch.as_tensor(np.asarray(obs[cam]))
        b, t, h, w, c = img.shape
        feats[cam] = O.trunk_forward(p, cam, img.permute(0, 2, 3, 1, 4).reshape(b, h, w, t * c), dtype)
    masks = dropout_masks if dropout_masks is not None else O._dropout_masks(drop_key, cams, B)
    masks = {c: torch.as_tensor(np.asarray(m)).bool() for c, m in masks.items()}
    mu, sd = bc_forward(full, cams, feats, torch.as_tensor(np.asarray(obs["state"])).to(dtype), masks, std_min, std_max)
    z = (actions - mu) / sd
    logp = (-0.5 * z * z - torch.log(sd) - 0.5 * math.log(2 * math.pi)).sum(-1)
    loss = -logp.mean()
    mse = ((mu - actions) ** 2).sum(-1).mean()
    gs = torch.autograd.grad(loss, list(train.values()), allow_unused=True)
    grads = {kk: (torch.zeros_like(v) if g is None else g) for (kk, v), g in zip(train.items(), gs)}
    upd = O.adam_tx_update(grads, opt, lr)
    new_params = dict(p)
    for kk in train:
        new_params[kk] = p[kk] + upd[kk]
    return new_params, opt, new_rng, {"actor_loss": loss.item(), "mse": mse.item(), "_mu": mu.detach(), "_std": sd.detach()}, grads
