"""Checkpoint save / restore with the call signature the learner scripts use
(`flax.training.checkpoints.save_checkpoint(ckpt_dir, target, step=..., keep=...)`,
examples/async_drq_sim/async_drq_sim.py:303-307; `restore_checkpoint(ckpt_dir, target)`).

The payload is `TrainState.state_dict()`: {step, params, target_params, opt_states{actor,critic,temperature}, rng} as
nested dicts of NumPy arrays in the Flax tree layout - the same fields `JaxRLTrainState` holds
(reference common/common.py:108-114) - stored as one pickle per step (`checkpoint_<step>`), newest `keep` kept.
When flax is importable the state also registers with `flax.serialization`, so the reference's own
`checkpoints.save_checkpoint(path, agent.state, ...)` call works on it unchanged."""
from __future__ import annotations

import os
import pickle
import re


def _ckpts(ckpt_dir, prefix):
    pat = re.compile(re.escape(prefix) + r"(\d+)$")
    out = []
    if os.path.isdir(ckpt_dir):
        for f in os.listdir(ckpt_dir):
            m = pat.match(f)
            if m:
                out.append((int(m.group(1)), os.path.join(ckpt_dir, f)))
    return sorted(out)


def save_checkpoint(ckpt_dir, target, step, prefix="checkpoint_", keep=1, overwrite=False):
    os.makedirs(ckpt_dir, exist_ok=True)
    path = os.path.join(ckpt_dir, f"{prefix}{int(step)}")
    if os.path.exists(path) and not overwrite:
        raise FileExistsError(path)
    payload = target.state_dict() if hasattr(target, "state_dict") else target
    tmp = path + ".tmp"
    with open(tmp, "wb") as f:
        pickle.dump(payload, f, protocol=4)
    os.replace(tmp, path)
    for _, old in _ckpts(ckpt_dir, prefix)[:-keep]:
        os.remove(old)
    return path


def latest_checkpoint(ckpt_dir, prefix="checkpoint_"):
    c = _ckpts(ckpt_dir, prefix)
    return c[-1][1] if c else None


def restore_checkpoint(ckpt_dir, target, step=None, prefix="checkpoint_"):
    path = os.path.join(ckpt_dir, f"{prefix}{int(step)}") if step is not None else (
        ckpt_dir if os.path.isfile(ckpt_dir) else latest_checkpoint(ckpt_dir, prefix))
    if path is None or not os.path.exists(path):
        return target
    with open(path, "rb") as f:
        payload = pickle.load(f)
    if target is not None and hasattr(target, "load_state_dict"):
        return target.load_state_dict(payload)
    return payload
