"""DrQAgent on hand-written sm_100a kernels.

Mirrors the reference's `DrQAgent` (agents/continuous/drq.py:23-328): `create_drq`, `update_critics`,
`update_high_utd`, inheriting `update` / `sample_actions` from SACAgent.  The DrQ random shift is not a
separate pass: it is applied by the replay sampler kernel while it gathers the frames, keyed by the
same JAX key chain as `data_augmentation_fn` (drq.py:244-253,307-310; same offsets for every camera of
a sample, different keys for obs and next_obs).

Only `encoder_type="resnet-pretrained"` is implemented - the one encoder every SERL launcher uses; the
reference's "small" and "resnet" branches raise TypeError at the first forward.
"""
from __future__ import annotations

from typing import Iterable, Optional

import numpy as np

from ... import ops
from ...data.replay_buffer import BatchHandle
from ...engine import AgentConfig
from .sac import SACAgent, _check_architecture_kwargs, _leaf, register_pytree


class DrQAgent(SACAgent):
    @classmethod
    def create_drq(cls, seed: int, observations, actions, *, encoder_type: str = "small", use_proprio: bool = True,
                   image_keys: Iterable[str] = ("image",), discount: float = 0.95, critic_ensemble_size: int = 2,
                   critic_subsample_size: Optional[int] = None, temperature_init: float = 1.0, backup_entropy: bool = False,
                   soft_target_update_rate: float = 0.005, target_entropy: Optional[float] = None, policy_kwargs=None,
                   learning_rate: float = 3e-4, precision: str = "fp32", device=None, **kwargs):
        _check_architecture_kwargs(policy_kwargs, kwargs, pixel=True)
        if encoder_type != "resnet-pretrained":
            raise NotImplementedError(f"encoder_type={encoder_type!r}: only 'resnet-pretrained' is supported "
                                      "(the reference's 'small'/'resnet' paths are broken, SURVEY.md Appendix C.1)")
        if not use_proprio:
            raise NotImplementedError("use_proprio=False is not used by any SERL launcher")
        pk = policy_kwargs or {}
        image_keys = tuple(image_keys)
        st = np.asarray(observations["state"])
        img = np.asarray(observations[image_keys[0]])
        T = img.shape[-4] if img.ndim >= 4 else 1
        if T != 1:
            raise NotImplementedError("obs_horizon must be 1 (ChunkingWrapper(obs_horizon=1) in every SERL example)")
        hw = img.shape[-2]
        S = int(np.prod(st.shape[-2:])) if st.ndim >= 2 else int(st.shape[-1])
        A = int(np.asarray(actions).shape[-1])
        cfg = AgentConfig(cams=image_keys, state_in=S, action_dim=A, pixel=True, ensemble=critic_ensemble_size,
                          subsample=critic_subsample_size, discount=discount, tau=soft_target_update_rate,
                          target_entropy=(-A / 2 if target_entropy is None else target_entropy), backup_entropy=backup_entropy,
                          lr=(learning_rate,) * 3, warmup=(0, 0, 0),                 # drq.py:35-43: no warm-up
                          std_min=pk.get("std_min", 1e-5), std_max=pk.get("std_max", 10.0), image_hw=hw, precision=precision)
        agent = cls._build(seed, cfg, temperature_init, device, config_extra={"image_keys": image_keys})
        from ...utils.train_utils import load_resnet10_params
        return load_resnet10_params(agent, image_keys)                              # drq.py:237-240

    def update_critics(self, batch, *, pmap_axis: Optional[str] = None):
        """drq.py:296-328: unpack + augment + update{critic}."""
        B = batch.batch_size if isinstance(batch, BatchHandle) else int(np.asarray(_leaf(batch, "rewards")).shape[0])
        eng = self._engine(B)
        nets = frozenset({"critic"})

        def body(batch, graph_mode):
            ops.rng_schedule(self.state._rng, self._keys, True, True)    # split(rng,3) then update's split(rng,4)
            eng.launches += 1
            if getattr(eng, "fused", None) is not None and self.explicit_randomness is None:
                eng.fused.prefetch_rng(self._keys)
            with self._section("sample_crop"):
                self._load_batch(eng, batch, augment=True, keys=self._keys, graph_mode=graph_mode)
            with self._section("trunk"):
                self._features(eng)
            self._update_on_engine(eng, nets, pmap_axis, schedule_keys=False, want_info=False)

        self._run_step(self._graph_key(("update_critics", pmap_axis), batch), batch, body)
        info = self._info(eng, nets)
        del info["actor"], info["temperature"]
        return self, info

    def update_high_utd(self, batch, *, utd_ratio: int, pmap_axis: Optional[str] = None):
        """drq.py:255-294: augment once, then SACAgent.update_high_utd."""
        return super().update_high_utd(batch, utd_ratio=utd_ratio, pmap_axis=pmap_axis, _augment=True)


register_pytree(DrQAgent)
