"""SACAgent on hand-written sm_100a kernels.

Mirrors the public surface of the reference's `SACAgent` (agents/continuous/sac.py:21-596):
`create_states` / `create_pixels`-style construction, `update(batch, pmap_axis, networks_to_update)`,
`update_high_utd(batch, utd_ratio)`, `sample_actions(observations, seed, argmax)`, `state`, `config`,
`replace(state=...)`.  Calls return `(agent, info)` like the reference (the agent is updated in place:
its parameters live in HBM).  `info` leaves are 0-d device tensors; `float(x)` synchronises.

Semantics reproduced: ensemble subsample with replacement + min for the TD
target, mean over the ensemble in the actor loss, shared value head for the pixel agent, same key for
dropout and action sampling in `_compute_next_actions`, ALL three Adam txs tick on every `update`
(zero-gradient momentum drift), polyak over the whole tree after critic updates, JAX key chain.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Dict, FrozenSet, Optional, Sequence

import numpy as np
import torch

from ... import _lib as L
from ... import ops
from ...common.common import TrainState
from ...data.replay_buffer import BatchHandle
from ...engine import AgentConfig, Engine
from ...params import ParamStore, init_trainable, init_trunk, trainable_spec, trunk_spec

ALL_NETS = frozenset({"actor", "critic", "temperature"})


_HEADS_PDL = os.environ.get("SERL_HEADS_PDL", "0") not in ("", "0")
_SPLIT_ALLREDUCE = os.environ.get("SERL_SPLIT_ALLREDUCE", "0") not in ("", "0")


def _dist():
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            return dist
    except Exception:                                   # noqa: BLE001
        pass
    return None


class SACAgent:
    def __init__(self, cfg: AgentConfig, store: ParamStore, trunk, state: TrainState, config: dict, device):
        self._cfg, self._store, self._trunk, self.state, self.config, self.device = cfg, store, trunk, state, config, device
        self._engines: Dict[int, Engine] = {}
        self._keys = torch.zeros(2 * L.NUM_KEYS, dtype=torch.uint32, device=device)
        self._seed_key = torch.zeros(2, dtype=torch.uint32, device=device)
        self.data_parallel = False          # set True to all-reduce(mean) gradients + infos (reference: pmap_axis)
        self.explicit_randomness = None     # tests: dict with eps / dropout / subsample (and crop offsets)
        self.use_cuda_graphs = True         # replay the whole step as one CUDA graph from its 3rd identical call on
        self.section_events = None          # bench: list collecting (name, start event, end event) of eagerly launched steps
        # Cross-step pipeline (DrQ pixel agent, opt-in): the frozen encoder of step i+1 does not depend on the parameters step i
        # updates, so `update_critics` can run sampler + trunk of the NEXT sequential batch next to the heads / Adam of the current
        # one (drq.py::_update_critics_pipelined).  The next batch is then drawn one call early - like the reference iterator's
        # `queue_size=2` prefetch (data/replay_buffer.py:77-90) - and a call that does not continue the sequence falls back.
        self.pipeline_critic_steps = False
        self._pipe = None
        self._eng_pair: Dict[int, list] = {}
        self._keys_pair = [self._keys, torch.zeros_like(self._keys)]
        self._rng_look = torch.zeros(2, dtype=torch.uint32, device=device)
        self._pipe_stream = None
        self._last_engine = None
        self._graphs = {}
        self._graphs_version = store.version   # captured graphs bake parameter-derived state (packed trunk weights, stem sign mask)
        self._launch_adj = 0                # graph capture / replay correction of the library's launch counter

    # ---- construction (sac.py:322-400,486-542) ------------------------------------------------------
    @classmethod
    def _build(cls, seed: int, cfg: AgentConfig, temperature_init: float, device, in_channels: int = 3, config_extra=None):
        L.load()
        device = torch.device(device if device is not None else "cuda")
        L.require_cuda(device)
        rng = np.random.default_rng(seed)
        spec = trainable_spec(cfg.cams, cfg.state_in, cfg.action_dim, cfg.ensemble, cfg.pixel)
        store = ParamStore(spec, device)
        values = init_trainable(rng, spec, temperature_init)
        store.load(store.params, values)
        store.target.copy_(store.params)                               # target_params=params (sac.py:369)
        trunk = {}
        if cfg.pixel:
            for cam in cfg.cams:
                w = init_trunk(rng, in_channels)
                trunk[cam] = {k: torch.as_tensor(v).to(device).contiguous() for k, v in w.items()}
        # rng, init_rng = split(PRNGKey(seed)); rng, create_rng = split(rng)  (sac.py:360,368): state.rng = create_rng
        key = np.array([(seed >> 32) & 0xFFFFFFFF, seed & 0xFFFFFFFF], dtype=np.uint32)
        key = _host_split(key, 2)[0]
        create = _host_split(key, 2)[1]
        rng_dev = torch.zeros(2, dtype=torch.uint32, device=device)
        state = TrainState(store, trunk, rng_dev)
        state.replace(rng=create)
        config = dict(critic_ensemble_size=cfg.ensemble, critic_subsample_size=cfg.subsample, discount=cfg.discount,
                      soft_target_update_rate=cfg.tau, target_entropy=cfg.target_entropy, backup_entropy=cfg.backup_entropy)
        config.update(config_extra or {})
        return cls(cfg, store, trunk, state, config, device)

    @classmethod
    def create_states(cls, seed: int, observations, actions, *, discount=0.95, critic_ensemble_size=2,
                      critic_subsample_size=None, temperature_init=1.0, backup_entropy=False, soft_target_update_rate=0.005,
                      target_entropy=None, policy_kwargs=None, actor_warmup=2000, critic_warmup=2000, learning_rate=3e-4,
                      device=None, **kwargs):
        """State-observation agent (sac.py:486-542).  Optimizer defaults follow SACAgent.create (:333-343):
        2000-step linear warm-up for actor and critic."""
        _check_architecture_kwargs(policy_kwargs, kwargs, pixel=False)
        pk = policy_kwargs or {}
        S = int(np.asarray(observations).shape[-1])
        A = int(np.asarray(actions).shape[-1])
        cfg = AgentConfig(cams=(), state_in=S, action_dim=A, pixel=False, ensemble=critic_ensemble_size,
                          subsample=critic_subsample_size, discount=discount, tau=soft_target_update_rate,
                          target_entropy=(-A / 2 if target_entropy is None else target_entropy), backup_entropy=backup_entropy,
                          lr=(learning_rate,) * 3, warmup=(critic_warmup, actor_warmup, 0),
                          std_min=pk.get("std_min", 1e-5), std_max=pk.get("std_max", 10.0))
        return cls._build(seed, cfg, temperature_init, device)

    def replace(self, **kw):
        if "state" in kw:
            self.state = kw.pop("state")
            self.invalidate_graphs()
        if kw:
            raise TypeError(f"replace: unknown fields {sorted(kw)}")
        return self

    def invalidate_graphs(self):
        """Drops every captured CUDA graph (and the packed 16-bit trunk weights derived from the fp32 ones): called when
        parameters were written from outside the step (`state.replace(params=...)`, checkpoint restore)."""
        self._graphs.clear()
        self._pipe = None
        self._graphs_version = self._store.version
        for eng in list(self._engines.values()) + [e for pair in self._eng_pair.values() for e in pair]:
            eng.__dict__.pop("_tc_weights", None)

    # ---- engines ---------------------------------------------------------------------------------
    def _engine(self, B: int) -> Engine:
        if B not in self._engines:
            self._engines[B] = Engine(self._cfg, self._store, self._trunk, B, self.device)
        return self._engines[B]

    @property
    def kernel_launches(self) -> int:
        """Kernels of libserl_b200 executed so far in this process: the library's own launch counter, minus launches that
        were only recorded during graph capture, plus the recorded count for every replay."""
        return L.launch_count() + self._launch_adj

    # ---- CUDA graphs: the ~150 launches of a step are captured once and replayed -------------------------
    def _graph_key(self, tag, batch):
        """Batches that can be replayed: lazy handles whose index draw reads the ring's device-resident counters."""
        if not self.use_cuda_graphs or self.explicit_randomness is not None or not isinstance(batch, BatchHandle):
            return None
        if torch.device(self.device).type != "cuda":            # host-logic dry runs (tests) have nothing to capture
            return None
        if any(p.get("indx") is not None for p in batch.parts):
            return None
        return (tag, batch.batch_size, tuple((id(p["ring"]), p["batch"]) for p in batch.parts))

    def _run_step(self, key, batch, body):
        """body(batch, graph_mode) enqueues one step.  1st call with a key: eager (warm-up: lazy allocations, function
        attributes); 2nd: capture + replay; later: replay only."""
        if self._graphs_version != self._store.version:          # TrainState.replace(params=...) since the last capture
            self.invalidate_graphs()
        self._pipe = None                                        # any step outside the pipelined path consumes the key chain: prefetch is stale
        self._keys = self._keys_pair[0]
        if key is None:
            return body(batch, False)
        entry = self._graphs.get(key)
        if entry is None:
            self._graphs[key] = "warm"
            return body(batch, False)
        for p in batch.parts:                                    # device draw counter := this handle's step
            ring = p["ring"]
            if ring._dev_step_mirror != p["step"]:
                ring.step_dev.fill_(p["step"])
            ring._dev_step_mirror = p["step"] + 1
        if entry == "warm":
            g = torch.cuda.CUDAGraph()
            l0, s0, c0 = {b: e.launches for b, e in self._engines.items()}, self.state.step, L.launch_count()
            # thread-local capture mode + the ring locks: a DataStore insert thread must not enqueue its flush (an H2D copy
            # on another stream) into - or invalidate - this capture
            import contextlib
            with contextlib.ExitStack() as stack:
                for p in batch.parts:
                    lock = getattr(p["ring"], "_lock", None)
                    if lock is not None:
                        stack.enter_context(lock)
                with torch.cuda.graph(g, capture_error_mode="thread_local"):
                    body(batch, True)
            recorded = L.launch_count() - c0
            self._launch_adj -= recorded                              # recorded, not executed
            entry = (g, {b: e.launches - l0.get(b, 0) for b, e in self._engines.items()}, self.state.step - s0, recorded)
            self._graphs[key] = entry
            self.state.step = s0
            for b, e in self._engines.items():
                e.launches = l0.get(b, e.launches)
        g, launches, steps, recorded = entry
        g.replay()
        self._launch_adj += recorded
        for b, n in launches.items():
            self._engines[b].launches += n
        self.state.step += steps
        return None

    # ---- batch ingestion -------------------------------------------------------------------------------
    def _load_batch(self, eng: Engine, batch, *, augment: bool, keys, graph_mode: bool = False) -> None:
        """Fills the engine's batch buffers.  Pixel agents: obs crops to pix rows [0,B), next crops to [B,2B)."""
        cfg, B = self._cfg, eng.B
        if not isinstance(batch, BatchHandle):
            batch = self._handle_from_dict(batch)
        if batch.batch_size != B:
            raise ValueError(f"batch size {batch.batch_size} != engine batch {B}")
        out = L.BatchOut()
        if cfg.pixel:
            hw = cfg.image_hw
            for j, cam in enumerate(cfg.cams):
                out.obs_pix[j] = eng.pix[cam].data_ptr()
                out.next_pix[j] = eng.pix[cam].data_ptr() + B * hw * hw * 3
            out.off_obs, out.off_next = eng.off[0].data_ptr(), eng.off[1].data_ptr()
        out.obs_state, out.next_state, out.actions = eng.state_o.data_ptr(), eng.state_n.data_ptr(), eng.actions.data_ptr()
        out.rewards, out.masks, out.dones = eng.rewards.data_ptr(), eng.masks.data_ptr(), eng.dones.data_ptr()
        out.idx, out.status = eng.idx.data_ptr(), eng.status.data_ptr()
        expl = None
        if self.explicit_randomness is not None and "crop" in self.explicit_randomness:
            expl = tuple(torch.as_tensor(np.asarray(o), dtype=torch.int32, device=self.device).contiguous()
                         for o in self.explicit_randomness["crop"])
        elif not augment or not cfg.pixel:
            ident = torch.full((B, 2), 4, dtype=torch.int32, device=self.device)
            expl = (ident, ident)
        row = 0
        # RLPD (concat_batches of an online and a demo handle): the parts gather disjoint output rows from different rings, so the
        # second part's launch runs on side stream 0 next to the first (each launch alone is one partial wave of CTAs: latency-bound)
        side = eng.side[0] if (graph_mode and len(batch.parts) == 2 and batch.parts[0]["ring"] is not batch.parts[1]["ring"]) else None
        for pi, part in enumerate(batch.parts):
            ring = part["ring"]
            if cfg.pixel and ring.T != 1:
                raise NotImplementedError("the trunk kernels take one frame per observation (obs_horizon=1), like every SERL example")
            on_side = side is not None and pi == 1
            if on_side:
                side.fork()
                side.__enter__()
            try:
                ring.launch_sample(part, out, crop_total=B, out_row_offset=row, key_obs=ops.key_ptr(keys, L.KEY_CROP_OBS),
                                   key_next=ops.key_ptr(keys, L.KEY_CROP_NEXT), explicit_off=expl,
                                   step_dev=ring.step_dev if graph_mode else None, record_event=not graph_mode)
                eng.launches += 1
                if graph_mode:
                    ops.counter_add(ring.step_dev, 1)
                    eng.launches += 1
            finally:
                if on_side:
                    side.__exit__(None, None, None)
            row += part["batch"]
        if side is not None:
            side.join()

    def _handle_from_dict(self, batch: dict) -> BatchHandle:
        """Host / device dict in the reference layout -> a temporary HBM ring + explicit indices."""
        from ...data.replay_buffer import DeviceRing
        cfg = self._cfg
        dev = self.device
        t = lambda x, dt=torch.float32: torch.as_tensor(np.asarray(x) if not isinstance(x, torch.Tensor) else x).to(dev, dt)
        B = int(t(batch["rewards"]).shape[0])
        if cfg.pixel:
            obs, nobs = batch["observations"], batch["next_observations"]
            T = 1
            ring = DeviceRing(B * 2, cfg.cams, (cfg.image_hw, cfg.image_hw, 3), T, cfg.state_in, cfg.action_dim, device=dev, seed=0)
            for cam in cfg.cams:
                pix = t(obs[cam], torch.uint8)
                packed = pix if cam not in nobs else torch.cat([pix, t(nobs[cam], torch.uint8)[:, -1:]], dim=1)
                if packed.shape[1] != 2:
                    raise NotImplementedError("dict batches: obs_horizon must be 1")
                ring.frames[cam].copy_(packed.reshape(B * 2, *packed.shape[2:]))
            sl = slice(1, None, 2)
            ring.state[sl] = t(obs["state"]).reshape(B, -1)
            ring.next_state[sl] = t(nobs["state"]).reshape(B, -1)
            idx = torch.arange(B, device=dev, dtype=torch.int32) * 2 + 1
        else:
            ring = DeviceRing(B, (), (1, 1, 1), 1, cfg.state_in, cfg.action_dim, device=dev, seed=0)
            sl = slice(None)
            ring.state[sl] = t(batch["observations"]).reshape(B, -1)
            ring.next_state[sl] = t(batch["next_observations"]).reshape(B, -1)
            idx = torch.arange(B, device=dev, dtype=torch.int32)
        ring.actions[sl] = t(batch["actions"]).reshape(B, -1)
        ring.rewards[sl] = t(batch["rewards"])
        ring.masks[sl] = t(batch["masks"])
        ring.dones[sl] = t(batch["dones"], torch.uint8) if "dones" in batch else 0
        ring.valid.fill_(1)
        ring._size = ring._capacity
        ring.size_dev.fill_(ring._size)
        return BatchHandle([dict(ring=ring, seed=0, step=0, batch=B, indx=idx)], True)

    # ---- update (sac.py:243-299) ------------------------------------------------------------------------
    def _section(self, name):
        """Context manager timing one section of an EAGER step with CUDA events (bench.py's per-section timeline); no-op otherwise."""
        import contextlib
        if self.section_events is None or torch.cuda.is_current_stream_capturing():
            return contextlib.nullcontext()

        @contextlib.contextmanager
        def cm():
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            try:
                yield
            finally:
                b.record()
                self.section_events.append((name, a, b))
        return cm()

    def _features(self, eng: Engine):
        if not self._cfg.pixel:
            return
        cams = self._cfg.cams
        # cameras 1.. first, each on its own stream (fork / join = graph edges), then camera 0 on the current stream.  The fp32
        # build shares its activation scratch between cameras and stays serial.
        side = [c for c in cams if eng.cam_stream.get(c) is not None and self._cfg.precision != "fp32"]
        for cam in side:
            cs = eng.cam_stream[cam]
            cs.fork()
            with cs:
                eng.trunk_forward(cam, eng.pix[cam], eng.feats[cam])
        for cam in cams:
            if cam not in side:
                eng.trunk_forward(cam, eng.pix[cam], eng.feats[cam])
        for cam in side:
            eng.cam_stream[cam].join()

    def _dp(self, pmap_axis) -> bool:
        """ONE predicate for both halves of the data-parallel exchange (1/world pre-scaling in the loss kernels and the SUM
        all-reduce): the reference's `pmap_axis is not None`, or the `data_parallel` switch, in a multi-rank job."""
        return (pmap_axis is not None or self.data_parallel) and _dist() is not None

    def _allreduce(self, lo: int, hi: int):
        """jax.lax.pmean(grads_and_aux) (common.py:213-214) as ONE collective: the loss kernels scale gradients AND info
        scalars by 1/world, and the infos sit inside the flat gradient buffer next to the segment they belong to
        (params.py), so a single SUM all-reduce of [lo, hi) yields the mean of both."""
        _dist().all_reduce(self._store.grad[lo:hi], op=_dist().ReduceOp.SUM)

    def _update_on_engine(self, eng: Engine, nets: FrozenSet[str], pmap_axis=None, schedule_keys: bool = True, want_info: bool = True):
        assert nets.issubset(ALL_NETS), f"Invalid gradient steps: {nets}"
        if schedule_keys:
            ops.rng_schedule(self.state._rng, self._keys, False, True)
            eng.launches += 1
        expl = self.explicit_randomness
        st = self._store
        dp = self._dp(pmap_axis)
        gscale = 1.0 / _dist().get_world_size() if dp else 1.0
        at = "actor" in nets or "temperature" in nets
        # programmatic dependent launch for the ~40-launch heads chain only (SERL_HEADS_PDL=1): the next kernel's CTAs become
        # resident and run their set-up while the current one drains; the trunk's one-CTA-per-SM kernels gain nothing from it
        heads_pdl = _HEADS_PDL and torch.device(self.device).type == "cuda"
        if heads_pdl:
            L.call("serl_set_pdl", 1)
        # SERL_SPLIT_ALLREDUCE=1 (experimental, off: ONE collective per step by default): the critic-MLP gradients (+ the info scalars
        # right behind them in the flat buffer) are complete while the encoder backward still runs - their all-reduce goes out on the
        # weight-gradient side stream and overlaps it, the encoder segment follows at the end.  Measured on 2 GPUs (trunk-bound per
        # rank, the heads chain is hidden by the step pipeline): 1186 vs 1200 steps/s, i.e. the extra collective costs more than it
        # hides there; not measured at 8 GPUs, where the heads chain + all-reduce is what bounds the step.
        early = None
        fused = getattr(eng, "fused", None)
        if dp and fused is not None and nets == frozenset({"critic"}) and _SPLIT_ALLREDUCE:
            c0 = st.leaf["modules_critic/network/Dense_0/kernel"].offset
            early = (c0, st.info_off + 4)
            fused.early_allreduce = lambda: self._allreduce(*early)
        with self._section("heads"):
            if "critic" in nets:
                eng.critic_loss_and_grads(self._keys, grad_scale=gscale, explicit=expl)
            if at:
                # any subset is legal (sac.py:270-277): a network that is not updated contributes a zero gradient, its tx still ticks
                eng.actor_temp_loss_and_grads(self._keys, grad_scale=gscale, explicit=expl, do_actor="actor" in nets,
                                              do_temperature="temperature" in nets)
        if heads_pdl:
            L.call("serl_set_pdl", int(os.environ.get("SERL_PDL", "0") not in ("", "0")))
        if dp and nets:
            # [group 0 | critic infos] and/or [actor, temperature infos | groups 1, 2 | aux]: one contiguous range either way
            with self._section("allreduce"):
                if early is not None:
                    fused.early_allreduce = None
                    self._allreduce(0, early[0])
                else:
                    self._allreduce(0 if "critic" in nets else st.info_off + 4, st.n if at else st.info_off + 4)
        with self._section("adam_polyak"):
            eng.optimizer_step([int("critic" in nets), int("actor" in nets), int("temperature" in nets)], polyak="critic" in nets)
        self.state.step += 1
        return self._info(eng, nets) if want_info else None

    def _info(self, eng: Engine, nets) -> dict:
        snap = torch.cat([eng.info[:12], eng.lr_info])
        info = {"critic": {}, "actor": {}, "temperature": {}}
        if "critic" in nets:
            info["critic"] = {"critic_loss": snap[0], "predicted_qs": snap[1], "target_qs": snap[2]}
        if "actor" in nets:
            info["actor"] = {"actor_loss": snap[4], "temperature": snap[5], "entropy": snap[6]}
        if "temperature" in nets:
            info["temperature"] = {"temperature_loss": snap[8]}
        info["critic_lr"], info["actor_lr"], info["temperature_lr"] = snap[12], snap[13], snap[14]   # sac.py:292-297
        return info

    def update(self, batch, *, pmap_axis: Optional[str] = None, networks_to_update: FrozenSet[str] = ALL_NETS):
        """One gradient step on all (or a subset of) the networks (sac.py:243-299)."""
        nets = frozenset(networks_to_update)
        B = batch.batch_size if isinstance(batch, BatchHandle) else int(np.asarray(_leaf(batch, "rewards")).shape[0])
        eng = self._engine(B)

        def body(batch, graph_mode):
            ops.rng_schedule(self.state._rng, self._keys, False, True)
            eng.launches += 1
            self._load_batch(eng, batch, augment=False, keys=self._keys, graph_mode=graph_mode)
            self._features(eng)
            self._update_on_engine(eng, nets, pmap_axis, schedule_keys=False, want_info=False)

        self._run_step(self._graph_key(("update", tuple(sorted(nets)), pmap_axis), batch), batch, body)
        return self, self._info(eng, nets)

    def _check(self, eng: Engine):
        pass   # draw failures are surfaced lazily by check_status() to avoid a sync per step

    def check_status(self):
        for eng in list(self._engines.values()) + [pair[1] for pair in self._eng_pair.values()]:
            if int(eng.status.item()):
                raise L.SerlError("replay draw failed: no valid slot within the redraw budget")
            if getattr(eng, "fused", None) is not None:
                eng.fused.check_error()

    def update_high_utd(self, batch, *, utd_ratio: int, pmap_axis: Optional[str] = None, _augment: bool = False):
        """sac.py:544-596: utd_ratio critic updates on consecutive minibatches, then one actor+temperature update
        on the full batch."""
        self._pipe, self._keys = None, self._keys_pair[0]          # consumes the key chain: a prefetched next batch is stale
        B = batch.batch_size if isinstance(batch, BatchHandle) else int(np.asarray(_leaf(batch, "rewards")).shape[0])
        assert B % utd_ratio == 0, f"Batch size {B} must be divisible by UTD ratio {utd_ratio}"
        full = self._engine(B)
        mb = B // utd_ratio
        if utd_ratio == 1:                                         # the DrQ learner's case: one graph for the whole call
            def body(batch, graph_mode):
                if _augment:
                    ops.rng_schedule(self.state._rng, self._keys, True, False)      # drq.py:279
                    full.launches += 1
                self._load_batch(full, batch, augment=_augment, keys=self._keys, graph_mode=graph_mode)
                self._features(full)
                self._update_on_engine(full, frozenset({"critic"}), pmap_axis, want_info=False)
                full.info_hist.copy_(full.info)                                      # critic infos of the scan step
                self._update_on_engine(full, frozenset({"actor", "temperature"}), pmap_axis, want_info=False)

            self._run_step(self._graph_key(("high_utd", _augment, pmap_axis), batch), batch, body)
            at = self._info(full, frozenset({"actor", "temperature"}))
            snap = full.info_hist.clone()
            info = {"critic": {"critic_loss": snap[0], "predicted_qs": snap[1], "target_qs": snap[2]},
                    "actor": at["actor"], "temperature": at["temperature"]}
            for k in ("critic_lr", "actor_lr", "temperature_lr"):
                info[k] = at[k]
            return self, info
        if _augment:
            ops.rng_schedule(self.state._rng, self._keys, True, False)          # drq.py:279
            full.launches += 1
        self._load_batch(full, batch, augment=_augment, keys=self._keys)
        self._features(full)
        crit_infos = []
        for i in range(utd_ratio):
            eng = self._minibatch_engine(full, i, mb)
            crit_infos.append(self._update_on_engine(eng, frozenset({"critic"}), pmap_axis))
        at = self._update_on_engine(full, frozenset({"actor", "temperature"}), pmap_axis)
        crit = {k: torch.stack([c["critic"][k] for c in crit_infos]).mean() for k in crit_infos[0]["critic"]}
        info = {"critic": crit, "actor": at["actor"], "temperature": at["temperature"]}
        for k in ("critic_lr", "actor_lr", "temperature_lr"):
            info[k] = at[k]
        return self, info

    def _minibatch_engine(self, full: Engine, i: int, mb: int) -> Engine:
        eng = self._engine(mb)
        lo, hi, B = i * mb, (i + 1) * mb, full.B
        for name in ("state_o", "state_n", "actions", "rewards", "masks"):
            getattr(eng, name).copy_(getattr(full, name)[lo:hi])
        if self._cfg.pixel:
            for cam in self._cfg.cams:
                eng.feats[cam][:mb].copy_(full.feats[cam][lo:hi])
                eng.feats[cam][mb:].copy_(full.feats[cam][B + lo:B + hi])
        return eng

    # ---- sample_actions (sac.py:301-320) ---------------------------------------------------------------
    def sample_actions(self, observations, *, seed=None, argmax: bool = False, return_device: bool = False, **kwargs):
        cfg, dev = self._cfg, self.device
        if argmax:
            assert seed is None, "Cannot specify seed when sampling deterministically"
        if cfg.pixel:
            st = np.asarray(observations["state"])
            unbatched = st.ndim == 2                                        # (T,S)
            B = 1 if unbatched else st.shape[0]
            eng = self._engine(B)
            for cam in cfg.cams:
                img = torch.as_tensor(np.asarray(observations[cam])).to(dev)
                eng.pix[cam][:B].copy_(img.reshape(B, cfg.image_hw, cfg.image_hw, 3))
                eng.trunk_forward(cam, eng.pix[cam][:B], eng.feats[cam])
            eng.state_o.copy_(torch.as_tensor(st, dtype=torch.float32).reshape(B, -1))
        else:
            st = np.asarray(observations)
            unbatched = st.ndim == 1
            B = 1 if unbatched else st.shape[0]
            eng = self._engine(B)
            eng.state_o.copy_(torch.as_tensor(st, dtype=torch.float32).reshape(B, -1))
        eng.encode(self._store.params, slice(0, B), eng.state_o, eng.Xp, eng.F, None, save=False)   # train=False: no dropout
        eng.policy_forward(self._store.params, eng.Xp, save=False)
        A = cfg.action_dim
        if not argmax:
            key = np.ascontiguousarray(np.asarray(seed), dtype=np.uint32).reshape(2)
            self._seed_key.copy_(torch.from_numpy(key.view(np.int32)).view(torch.uint32))
            ops.normal_fill(self._seed_key.data_ptr(), eng.eps, B * A)
        ops.tanh_gaussian_fwd(eng.mu, eng.ls, eng.eps, cfg.std_min, cfg.std_max, eng.act_scratch.data_ptr(), A, None, None, None,
                              B, A, deterministic=argmax)
        out = eng.act_scratch.clone()
        out = out[0] if unbatched else out
        return out if return_device else out.cpu().numpy()


_LAUNCHER_NET_KWARGS = {"activations": "tanh", "use_layer_norm": True, "hidden_dims": [256, 256]}     # utils/launcher.py:61-66,95-104


def _check_architecture_kwargs(policy_kwargs, extra, pixel):
    """The kernels implement ONE architecture - the one every SERL launcher builds (utils/launcher.py:50-116): tanh MLPs
    [256, 256] with LayerNorm, tanh-squashed Gaussian policy with "exp" std parameterisation, shared encoder.  The
    reference's own defaults when these kwargs are omitted differ (swish, no LayerNorm, "uniform" std; sac.py:402-411,
    drq.py:113-131), so silently accepting other settings would build a different model than the caller asked for."""
    pk = dict(policy_kwargs or {})
    if pk.get("std_parameterization", "exp") != "exp" or not pk.get("tanh_squash_distribution", True) or pk.get("fixed_std") is not None:
        raise NotImplementedError(f"policy_kwargs={pk}: only the launcher's tanh-squashed 'exp' std parameterisation is implemented")
    for name in ("critic_network_kwargs", "policy_network_kwargs"):
        nk = extra.pop(name, None)
        if nk is None:
            continue
        act = nk.get("activations", "tanh")
        act = getattr(act, "__name__", act)
        if act != "tanh" or not nk.get("use_layer_norm", True) or list(nk.get("hidden_dims", [256, 256])) != [256, 256] \
                or nk.get("dropout_rate") not in (None, 0, 0.0):
            raise NotImplementedError(f"{name}={nk}: only {_LAUNCHER_NET_KWARGS} (utils/launcher.py) is implemented")
    if extra.pop("shared_encoder", True) is not True:
        raise NotImplementedError("shared_encoder=False is not implemented (every SERL launcher shares the encoder)")
    for k in ("actor_optimizer_kwargs", "critic_optimizer_kwargs", "temperature_optimizer_kwargs", "image_keys", "augmentation_function"):
        if extra.get(k) not in (None, {}):
            if k.endswith("optimizer_kwargs") and set(extra[k]) <= {"learning_rate"} and extra[k].get("learning_rate", 3e-4) == 3e-4:
                extra.pop(k)
                continue
            raise NotImplementedError(f"{k}={extra[k]!r} is not supported; use learning_rate=")
        extra.pop(k, None)
    if extra:
        raise TypeError(f"unexpected keyword arguments {sorted(extra)}")


def register_pytree(cls):
    """The learner scripts treat the agent as a JAX pytree: `jax.device_put(jax.tree_map(jnp.array, agent), sharding)`
    (examples/async_drq_sim/async_drq_sim.py:347-349) and `jax.block_until_ready(agent)` (:296).  This agent's arrays live in
    HBM behind the C-ABI, so it registers as a LEAF-LESS pytree node (the whole object travels as static aux data): both calls
    become identity operations.  No-op when jax is not importable."""
    try:
        from jax import tree_util
    except Exception:                                   # noqa: BLE001
        return False
    try:
        tree_util.register_pytree_node(cls, lambda a: ((), a), lambda aux, _children: aux)
    except ValueError:                                  # already registered
        pass
    return True


register_pytree(SACAgent)


def _host_split(key: np.ndarray, n: int) -> np.ndarray:
    """jax.random.split on the host through the library's host mirror of the device PRNG."""
    key = np.ascontiguousarray(key, dtype=np.uint32)
    out = np.zeros((n, 2), dtype=np.uint32)
    L.call("serl_host_threefry_split", key.ctypes.data, n, out.ctypes.data)
    return out


def _leaf(batch, key):
    v = batch[key]
    return v.cpu() if isinstance(v, torch.Tensor) else v
