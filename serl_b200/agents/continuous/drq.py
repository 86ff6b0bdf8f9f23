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

import os
from typing import Iterable, Optional

import numpy as np
import torch

from ... import _lib as L
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
        if (self.pipeline_critic_steps and self._cfg.pixel and self.section_events is None
                and self._graph_key(("update_critics", pmap_axis), batch) is not None):
            return self._update_critics_pipelined(batch, B, pmap_axis)
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

    # ---- cross-step pipeline: heads / Adam of step i next to sampler + frozen trunk of step i+1 ---------------------------
    def _update_critics_pipelined(self, batch: BatchHandle, B: int, pmap_axis):
        """`update_critics` for a batch handle that continues a sequence of handles (same rings, step + 1): the step's own
        sampler + trunk results were produced by the PREVIOUS call on the other engine of a ping-pong pair, and this call
        produces the next step's while its heads, all-reduce and Adam run (kind "P").  A call that does not continue the
        sequence runs its own front end first (kind "W").  Same kernels, same key chain, same results as the serial path as
        long as nothing is inserted between a prefetch and its use (then the prefetched draw simply predates the insert, as
        with the reference iterator's queue)."""
        import contextlib
        from ...engine import Engine
        nets = frozenset({"critic"})
        if self._graphs_version != self._store.version:
            self.invalidate_graphs()
        if B not in self._eng_pair:
            self._eng_pair[B] = [self._engine(B), Engine(self._cfg, self._store, self._trunk, B, self.device)]
        if self._pipe_stream is None:
            self._pipe_stream = L.new_side_stream(torch.device(self.device), True)
            # SERL_HEADS_PRIORITY=1: the heads chain on a high-priority stream (its CTAs are placed before the trunk's whenever an SM
            # frees up).  Measured 731 vs 745 steps/s at batch 256: the trunk's balanced grids already leave 20 SMs free, and what
            # limits the overlap there is that the bandwidth-type head kernels (SLE, Adam, reductions) get ~20 SMs - off by default.
            self._heads_stream = L.new_side_stream(torch.device(self.device), os.environ.get("SERL_HEADS_PRIORITY", "0") != "0", priority=-5)
        pair, Q, H = self._eng_pair[B], self._pipe_stream, self._heads_stream
        sig = (B, pmap_axis, tuple((id(p["ring"]), p["batch"], p["seed"]) for p in batch.parts))
        steps = tuple(p["step"] for p in batch.parts)
        pipe = self._pipe
        hit = pipe is not None and pipe["sig"] == sig and pipe["steps"] == steps
        par = pipe["par"] if hit else 0
        kind = "P" if hit else "W"
        cur, nxt = pair[par], pair[1 - par]
        Kc, Kn = self._keys_pair[par], self._keys_pair[1 - par]
        nxt_handle = BatchHandle([dict(p, step=p["step"] + 1) for p in batch.parts], batch.pack)

        def body(graph_mode):
            self._keys = Kc
            if kind == "W":                                          # this step's own front end (cold start of the pipeline)
                ops.rng_schedule(self.state._rng, Kc, True, True)
                cur.launches += 1
                self._rng_look.copy_(self.state._rng)
                if cur.fused is not None:
                    cur.fused.fill_rng_now(Kc)
                self._load_batch(cur, batch, augment=True, keys=Kc, graph_mode=graph_mode)
                self._features(cur)
            Q.fork()
            with Q:                                                  # front end of the NEXT step
                self.state._rng.copy_(self._rng_look)                # the key this step leaves behind (= what the serial path leaves)
                ops.rng_schedule(self._rng_look, Kn, True, True)
                nxt.launches += 1
                if nxt.fused is not None:
                    nxt.fused.fill_rng_now(Kn)
                self._load_batch(nxt, nxt_handle, augment=True, keys=Kn, graph_mode=graph_mode)
                self._features(nxt)
            H.fork()
            with H:
                self._update_on_engine(cur, nets, pmap_axis, schedule_keys=False, want_info=False)
            H.join()
            Q.join()

        gkey = ("pipe", kind, par, sig)
        entry = self._graphs.get(gkey) if self.use_cuda_graphs else "eager"
        if entry is None or entry == "eager":                         # first use of a variant: eager (lazy allocations), then capture
            if entry is None:
                self._graphs[gkey] = "warm"
            body(False)
        else:
            for p in batch.parts:                                     # device draw counters: W draws step then step + 1, P draws step + 1
                ring, need = p["ring"], p["step"] + (1 if kind == "P" else 0)
                if ring._dev_step_mirror != need:
                    ring.step_dev.fill_(need)
                ring._dev_step_mirror = need + (1 if kind == "P" else 2)
            if entry == "warm":
                g = torch.cuda.CUDAGraph()
                l0, s0, c0 = [e.launches for e in pair], self.state.step, L.launch_count()
                with contextlib.ExitStack() as stack:
                    for p in batch.parts:
                        lock = getattr(p["ring"], "_lock", None)
                        if lock is not None:
                            stack.enter_context(lock)
                    with torch.cuda.graph(g, capture_error_mode="thread_local"):
                        body(True)
                recorded = L.launch_count() - c0
                self._launch_adj -= recorded
                entry = (g, [e.launches - l for e, l in zip(pair, l0)], self.state.step - s0, recorded)
                self._graphs[gkey] = entry
                self.state.step = s0
                for e, l in zip(pair, l0):
                    e.launches = l
            g, launches, dsteps, recorded = entry
            g.replay()
            self._launch_adj += recorded
            for e, n in zip(pair, launches):
                e.launches += n
            self.state.step += dsteps
        self._keys = Kc
        self._pipe = dict(sig=sig, steps=tuple(s + 1 for s in steps), par=1 - par)
        self._last_engine = cur                                       # the engine whose buffers hold this step
        info = self._info(cur, nets)
        del info["actor"], info["temperature"]
        return self, info

    def update_high_utd(self, batch, *, utd_ratio: int, pmap_axis: Optional[str] = None):
        """drq.py:255-294: augment once, then SACAgent.update_high_utd."""
        return super().update_high_utd(batch, utd_ratio=utd_ratio, pmap_axis=pmap_axis, _augment=True)


register_pytree(DrQAgent)
