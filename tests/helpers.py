"""Shared helpers for the GPU parity tests (product objects <-> oracle objects)."""
from __future__ import annotations

import types

import numpy as np
import torch


class Box:
    def __init__(self, shape, dtype=np.float32):
        self.shape, self.dtype = tuple(shape), np.dtype(dtype)


class DictSpace:
    def __init__(self, spaces):
        self.spaces = dict(spaces)


def pixel_spaces(cams, hw=128, T=1, S=7, A=4):
    obs = DictSpace({**{c: Box((T, hw, hw, 3), np.uint8) for c in cams}, "state": Box((T, S))})
    return obs, Box((A,))


def fake_env(cams, hw=128, T=1, S=7, A=4):
    o, a = pixel_spaces(cams, hw, T, S, A)
    return types.SimpleNamespace(observation_space=o, action_space=a)


def random_transitions(rng, n, cams, hw=128, T=1, S=7, A=4, mean_ep=12):
    """Consecutive transitions share frames like a real episode (next_obs of t == obs of t+1)."""
    out, cur = [], None
    for _ in range(n):
        if cur is None:
            cur = {c: rng.integers(0, 256, (T, hw, hw, 3), dtype=np.uint8) for c in cams}
            cur["state"] = rng.standard_normal((T, S)).astype(np.float32)
        nxt = {c: np.concatenate([cur[c][1:], rng.integers(0, 256, (1, hw, hw, 3), dtype=np.uint8)]) for c in cams}
        nxt["state"] = rng.standard_normal((T, S)).astype(np.float32)
        done = bool(rng.random() < 1.0 / mean_ep)
        out.append(dict(observations=cur, next_observations=nxt, actions=rng.uniform(-1, 1, A).astype(np.float32),
                        rewards=np.float32(rng.random()), masks=np.float32(0.0 if done else 1.0), dones=done))
        cur = None if done else nxt
    return out


def to_numpy_tree(d):
    if isinstance(d, dict):
        return {k: to_numpy_tree(v) for k, v in d.items()}
    return d.detach().cpu().numpy() if isinstance(d, torch.Tensor) else np.asarray(d)


def oracle_state_from_agent(agent, dtype=torch.float64):
    from oracle.drq import OracleState
    from serl_b200.params import flatten
    st = agent.state
    params = {k: torch.as_tensor(np.asarray(v)) for k, v in flatten(st.params).items()}
    o = OracleState.create(params, st.rng, dtype)
    o.target_params = {k: torch.as_tensor(np.asarray(v)).to(dtype) for k, v in flatten(st.target_params).items()}
    os_ = st.opt_states
    for name in ("actor", "critic", "temperature"):
        o.opt[name]["count"] = os_[name]["count"]
        o.opt[name]["mu"].update({k: torch.as_tensor(v).to(dtype) for k, v in flatten(os_[name]["mu"]).items()})
        o.opt[name]["nu"].update({k: torch.as_tensor(v).to(dtype) for k, v in flatten(os_[name]["nu"]).items()})
    return o


def oracle_cfg_from_agent(agent):
    from oracle.drq import OracleConfig
    c = agent._cfg
    return OracleConfig(cams=tuple(c.cams), discount=c.discount, tau=c.tau, target_entropy=c.target_entropy,
                        ensemble=c.ensemble, subsample=c.subsample, backup_entropy=c.backup_entropy, lr=c.lr[0],
                        warmup={"critic": c.warmup[0], "actor": c.warmup[1], "temperature": c.warmup[2]}, pixel=c.pixel)


def rel_err(a, b):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-12))
