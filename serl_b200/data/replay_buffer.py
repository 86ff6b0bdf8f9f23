"""HBM-resident replay ring + lazy batch handles.

Mirrors the reference's `ReplayBuffer` (data/replay_buffer.py:40-90: ring storage, `insert`,
`__len__`, `get_iterator`) on top of a device ring: storage lives in HBM, inserts are staged in
pinned host memory and applied by a scatter kernel, and `sample` / the iterator return a
`BatchHandle` - indices are drawn, gathered and (for pixels) DrQ-shifted by ONE kernel when the agent
consumes the handle, so the reference's host gather + `jax.device_put` (replay_buffer.py:82-85)
disappears.  Index draws follow this repo's counter-based spec (oracle/replay.py::draw_indices).
"""
from __future__ import annotations

import ctypes as C
import threading
from typing import Dict, Iterable, List, Optional, Sequence

import numpy as np
import torch

from .. import _lib as L


def _space_shape(space):
    return tuple(space.shape)


def _is_dict_space(space):
    return hasattr(space, "spaces")


class BatchHandle:
    """A not-yet-materialised minibatch: (ring, seed, step, rows) parts, in concat order."""

    def __init__(self, parts: List[dict], pack_obs_and_next_obs: bool = True):
        self.parts = parts
        self.pack = pack_obs_and_next_obs
        self._dict = None

    @property
    def batch_size(self) -> int:
        return sum(p["batch"] for p in self.parts)

    def concat(self, other: "BatchHandle") -> "BatchHandle":
        return BatchHandle(self.parts + other.parts, self.pack)

    # dict-style access materialises an un-augmented copy in the reference's layout
    def to_dict(self) -> dict:
        if self._dict is None:
            outs = [p["ring"]._gather_dict(p, self.pack) for p in self.parts]
            self._dict = outs[0] if len(outs) == 1 else _cat_dicts(outs)
        return self._dict

    def __getitem__(self, k):
        return self.to_dict()[k]

    def keys(self):
        return self.to_dict().keys()


def _cat_dicts(ds):
    out = {}
    for k, v in ds[0].items():
        out[k] = _cat_dicts([d[k] for d in ds]) if isinstance(v, dict) else torch.cat([d[k] for d in ds], dim=0)
    return out


class DeviceRing:
    """Ring storage in HBM + host bookkeeping shared by both buffer flavours."""

    STAGE = 512

    def __init__(self, capacity: int, cams: Sequence[str], frame_shape, num_stack: int, state_dim: int, action_dim: int,
                 device=None, seed: Optional[int] = None):
        self.device = torch.device(device if device is not None else "cuda")
        L.require_cuda(self.device)
        L.load()
        self._capacity = int(capacity)
        self.cams = tuple(cams)
        self.frame_shape = tuple(frame_shape) if cams else (1, 1, 1)
        self.T, self.S, self.A = int(num_stack), int(state_dim), int(action_dim)
        dev, cap = self.device, self._capacity
        self.frames = {c: torch.zeros((cap, *self.frame_shape), dtype=torch.uint8, device=dev) for c in self.cams}
        f = lambda *s: torch.zeros(*s, dtype=torch.float32, device=dev)
        self.state, self.next_state = f(cap, self.T * self.S), f(cap, self.T * self.S)
        self.actions, self.rewards, self.masks = f(cap, self.A), f(cap), f(cap)
        self.dones = torch.zeros(cap, dtype=torch.uint8, device=dev)
        self.valid = torch.zeros(cap, dtype=torch.uint8, device=dev)
        self.size_dev = torch.zeros(1, dtype=torch.int32, device=dev)
        self.step_dev = torch.zeros(1, dtype=torch.int64, device=dev)     # graph-replay draw counter
        self._valid_host = np.zeros(cap, dtype=bool)
        self._size = 0
        self._insert_index = 0
        self._seed = int(seed) if seed is not None else int(np.random.SeedSequence().entropy % (1 << 63))
        self._draw_step = 0
        self._dev_step_mirror = 0                                         # host mirror of step_dev (graph replays)
        self._lock = threading.RLock()
        # Pinned staging, one interleaved record per staged slot write, behind a small header holding the validity changes:
        #   [touched slots int32 x 4n | touched values u8 x 4n | record 0 | record 1 | ...]
        # so a flush is ONE host->device copy of the used prefix, one scatter launch and one commit launch.  Two host
        # buffers alternate: the copy out of one overlaps the inserts into the other (no host sync per flush).
        n = self.STAGE
        fb = int(np.prod(self.frame_shape)) if self.cams else 0
        ns = self.T * self.S
        off, fields = 0, {}
        for c in self.cams:
            fields[("frames", c)] = (off, np.uint8, self.frame_shape); off += (fb + 15) // 16 * 16
        for name, cnt in (("state", ns), ("next_state", ns), ("actions", self.A), ("rewards", 1), ("masks", 1)):
            fields[name] = (off, np.float32, (cnt,)); off += 4 * cnt
        for name in ("dst", "src"):
            fields[name] = (off, np.int32, (1,)); off += 4
        for name in ("dones", "valid"):
            fields[name] = (off, np.uint8, (1,)); off += 1
        self._row_bytes = (off + 15) // 16 * 16
        self._hdr_bytes = (4 * n * 4 + 4 * n + 15) // 16 * 16
        self._fields = fields
        total = self._hdr_bytes + n * self._row_bytes
        self._stage_host = [L.pin(torch.zeros(total, dtype=torch.uint8)) for _ in range(2)]
        self._stage_dev = torch.empty(total, dtype=torch.uint8, device=dev)
        self._stage_evt = [None, None]                                   # copy-out completion of each host buffer
        self._cur = 0
        self._stn = [self._stage_views(b.numpy()) for b in self._stage_host]
        self._n_pending = 0
        self._pending_dst = set()
        self._touched = set()
        self._sample_evt = None
        self.h2d_bytes = 0

    def _stage_views(self, base: np.ndarray) -> dict:
        """Strided numpy views of one staging buffer: field -> (STAGE, ...) array whose row k lives in record k."""
        n, rb, hb = self.STAGE, self._row_bytes, self._hdr_bytes
        out = {"touched": base[:4 * n * 4].view(np.int32), "touched_val": base[4 * n * 4:4 * n * 4 + 4 * n], "frames": {}}
        for key, (off, dt, shape) in self._fields.items():
            item = np.dtype(dt).itemsize
            inner = tuple(int(np.prod(shape[i + 1:])) * item for i in range(len(shape)))
            v = np.ndarray((n, *shape), dtype=dt, buffer=base, offset=hb + off, strides=(rb, *inner))
            if isinstance(key, tuple):
                out["frames"][key[1]] = v
            else:
                out[key] = v[:, 0] if shape == (1,) else v
        return out

    # ---- reference API ---------------------------------------------------------------------------
    def __len__(self) -> int:
        return self._size

    def seed(self, seed: Optional[int] = None):
        if seed is not None:
            self._seed = int(seed)
        return [self._seed]

    # ---- staging ---------------------------------------------------------------------------------
    def _stage_write(self, dst: int, *, src_slot: int = -1, frames=None, state=None, next_state=None, action=None,
                     reward=0.0, mask=0.0, done=False, valid=False):
        """Queue one slot write (replay_buffer.py:71-75 semantics for the slot at `dst`)."""
        if (self._n_pending == self.STAGE or dst in self._pending_dst or (src_slot >= 0 and src_slot in self._pending_dst)
                or len(self._touched) > 3 * self.STAGE):
            self.flush()
        k = self._n_pending
        if k == 0 and self._stage_evt[self._cur] is not None:   # this buffer's previous copy-out (two flushes ago) must be over
            self._stage_evt[self._cur].synchronize()
            self._stage_evt[self._cur] = None
        st = self._stn[self._cur]
        st["dst"][k], st["src"][k] = dst, src_slot
        if src_slot < 0:
            for c in self.cams:
                st["frames"][c][k] = frames[c]
            st["state"][k] = np.asarray(state, np.float32).reshape(-1)
            st["next_state"][k] = np.asarray(next_state, np.float32).reshape(-1)
            st["actions"][k] = np.asarray(action, np.float32).reshape(-1)
            st["rewards"][k], st["masks"][k], st["dones"][k] = float(reward), float(mask), int(bool(done))
        st["valid"][k] = int(valid)
        self._valid_host[dst] = valid
        self._pending_dst.add(dst)
        self._touched.add(dst)
        self._n_pending = k + 1

    def _advance(self):
        self._insert_index = (self._insert_index + 1) % self._capacity
        self._size = min(self._size + 1, self._capacity)

    def _mark(self, slot: int, valid: bool):
        self._valid_host[slot] = valid
        self._touched.add(slot)

    def view(self) -> L.ReplayView:
        v = L.ReplayView()
        for j, c in enumerate(self.cams):
            v.frames[j] = self.frames[c].data_ptr()
        v.state, v.next_state, v.actions = self.state.data_ptr(), self.next_state.data_ptr(), self.actions.data_ptr()
        v.rewards, v.masks, v.dones, v.valid = (self.rewards.data_ptr(), self.masks.data_ptr(), self.dones.data_ptr(),
                                                self.valid.data_ptr())
        v.num_cams = len(self.cams)
        v.height, v.width, v.channels = self.frame_shape
        v.num_stack, v.state_dim, v.action_dim = self.T, self.S, self.A
        v.capacity, v.size = self._capacity, self._size
        return v

    def flush(self):
        """Apply staged slot writes + validity changes on the current stream."""
        with self._lock:
            n = self._n_pending
            stream_ptr = L.stream_ptr()
            if self._sample_evt is not None:
                self._sample_evt.make_current_stream_wait()   # never overwrite slots a sampling kernel still reads
            if n or self._touched:
                cur = self._cur
                if self._stage_evt[cur] is not None:          # only reachable when a flush carries validity changes alone
                    self._stage_evt[cur].synchronize()
                    self._stage_evt[cur] = None
                st, host, dv = self._stn[cur], self._stage_host[cur], self._stage_dev
                m = len(self._touched)
                if m:
                    slots = np.fromiter(self._touched, dtype=np.int32, count=m)
                    st["touched"][:m] = slots
                    st["touched_val"][:m] = self._valid_host[slots]
                nbytes = self._hdr_bytes + n * self._row_bytes
                dv[:nbytes].copy_(host[:nbytes], non_blocking=True)       # the ONE host->device copy of this flush
                self.h2d_bytes += nbytes
                base = dv.data_ptr() + self._hdr_bytes
                if n:
                    rq = L.ScatterRequest()
                    rq.n, rq.row_stride = n, self._row_bytes
                    f = self._fields
                    rq.dst_slot, rq.src_slot = base + f["dst"][0], base + f["src"][0]
                    for j, c in enumerate(self.cams):
                        rq.frames[j] = base + f[("frames", c)][0]
                    rq.state, rq.next_state, rq.actions = base + f["state"][0], base + f["next_state"][0], base + f["actions"][0]
                    rq.rewards, rq.masks, rq.dones, rq.valid = (base + f["rewards"][0], base + f["masks"][0], base + f["dones"][0],
                                                                base + f["valid"][0])
                    v = self.view()
                    L.call("serl_replay_scatter", C.byref(v), C.byref(rq), stream_ptr)
                # validity changes (applied after the slot writes, like the host ring logic orders them) + the new size
                L.call("serl_replay_commit", self.valid.data_ptr(), dv.data_ptr(), dv.data_ptr() + 4 * self.STAGE * 4, m,
                       self.size_dev.data_ptr(), self._size, stream_ptr)
                evt = L.new_event()
                evt.record()
                self._stage_evt[cur] = evt
                self._cur = cur ^ 1
            self._n_pending = 0
            self._pending_dst.clear()
            self._touched.clear()

    # ---- insert / sample (state-only flavour; the frame-dedup flavour overrides insert) -----------------
    def insert(self, data_dict: dict):
        """replay_buffer.py:71-75."""
        with self._lock:
            self._stage_write(self._insert_index, frames={}, state=data_dict["observations"], next_state=data_dict["next_observations"],
                              action=data_dict["actions"], reward=data_dict["rewards"], mask=data_dict["masks"], done=data_dict["dones"],
                              valid=True)
            self._advance()

    def sample(self, batch_size: int, keys: Optional[Iterable[str]] = None, indx=None, pack_obs_and_next_obs: bool = False) -> BatchHandle:
        with self._lock:
            self.flush()
            if self._size <= (self.T if self.cams else 0):
                raise L.SerlError(f"replay buffer holds {self._size} slots; cannot sample")
            part = dict(ring=self, seed=self._seed, step=self._draw_step, batch=int(batch_size),
                        indx=None if indx is None else torch.as_tensor(np.asarray(indx), dtype=torch.int32, device=self.device))
            self._draw_step += 1
            return BatchHandle([part], pack_obs_and_next_obs)

    def get_iterator(self, queue_size: int = 2, sample_args: dict = {}, device=None):
        """replay_buffer.py:77-90.  No prefetch queue is needed: handles are lazy and the data never leaves HBM."""
        while True:
            yield self.sample(**sample_args)

    # ---- kernel launch used by the agents ---------------------------------------------------------
    def launch_sample(self, part: dict, out: L.BatchOut, *, crop_total: int, out_row_offset: int, key_obs=None, key_next=None,
                      explicit_off=None, padding: int = 4, step_dev=None, record_event: bool = True):
        with self._lock:
            rq = L.SampleRequest()
            rq.seed, rq.step, rq.lane_offset, rq.batch = part["seed"], part["step"], 0, part["batch"]
            rq.step_dev = None if step_dev is None else step_dev.data_ptr()
            rq.size_dev = self.size_dev.data_ptr()
            rq.explicit_idx = None if part.get("indx") is None else part["indx"].data_ptr()
            rq.key_obs, rq.key_next = key_obs, key_next
            if explicit_off is not None:
                rq.explicit_off_obs, rq.explicit_off_next = explicit_off[0].data_ptr(), explicit_off[1].data_ptr()
            rq.crop_total, rq.out_row_offset, rq.padding = crop_total, out_row_offset, padding
            v = self.view()
            L.call("serl_replay_sample_crop", C.byref(v), C.byref(rq), C.byref(out), L.stream_ptr())
            if record_event:
                evt = L.new_event()
                evt.record()
                self._sample_evt = evt

    def _gather_dict(self, part: dict, pack: bool) -> dict:
        """Un-augmented materialisation in the reference's batch layout (memory_efficient_replay_buffer.py:126-164)."""
        B, dev, T = part["batch"], self.device, self.T
        e = lambda *s, dt=torch.float32: torch.empty(*s, dtype=dt, device=dev)
        obs_pix = {c: e(B, T, *self.frame_shape, dt=torch.uint8) for c in self.cams}
        next_pix = {c: e(B, T, *self.frame_shape, dt=torch.uint8) for c in self.cams}
        out = L.BatchOut()
        for j, c in enumerate(self.cams):
            out.obs_pix[j], out.next_pix[j] = obs_pix[c].data_ptr(), next_pix[c].data_ptr()
        st, nst, ac, rw, mk = e(B, T * self.S), e(B, T * self.S), e(B, self.A), e(B), e(B)
        dn = e(B, dt=torch.uint8)
        idx = e(B, dt=torch.int32)
        status = torch.zeros(1, dtype=torch.int32, device=dev)
        out.obs_state, out.next_state, out.actions, out.rewards, out.masks = (st.data_ptr(), nst.data_ptr(), ac.data_ptr(),
                                                                              rw.data_ptr(), mk.data_ptr())
        out.dones, out.idx, out.status = dn.data_ptr(), idx.data_ptr(), status.data_ptr()
        ident = torch.full((B * T, 2), 4, dtype=torch.int32, device=dev)          # centre offset = identity shift
        self.launch_sample(part, out, crop_total=B * T, out_row_offset=0, explicit_off=(ident, ident))
        if int(status.item()):
            raise L.SerlError("replay draw failed: no valid slot within the redraw budget")
        state_shape = (B, T, self.S) if self.cams else (B, self.S)
        obs = {"state": st.view(state_shape)} if self.cams else st.view(state_shape)
        nobs = {"state": nst.view(state_shape)} if self.cams else nst.view(state_shape)
        for c in self.cams:
            if pack:                                    # frames [idx-T .. idx]: obs frames then the newest next frame
                obs[c] = torch.cat([obs_pix[c], next_pix[c][:, -1:]], dim=1)
            else:
                obs[c], nobs[c] = obs_pix[c], next_pix[c]
        return {"observations": obs, "next_observations": nobs, "actions": ac, "rewards": rw, "masks": mk,
                "dones": dn.bool(), "_indices": idx}


class ReplayBuffer(DeviceRing):
    """State-observation ring (reference data/replay_buffer.py:40-75), storage in HBM."""

    def __init__(self, observation_space, action_space, capacity: int, next_observation_space=None, device=None, seed=None):
        if _is_dict_space(observation_space):
            raise TypeError("ReplayBuffer holds flat observations; use MemoryEfficientReplayBuffer for pixel dicts")
        S = int(np.prod(_space_shape(observation_space)))
        A = int(np.prod(_space_shape(action_space)))
        super().__init__(capacity, (), (1, 1, 1), 1, S, A, device=device, seed=seed)
