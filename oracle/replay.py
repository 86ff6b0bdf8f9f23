"""ORACLE (test infrastructure, not product): NumPy restatement of the reference's
frame-deduplicating replay ring and of this repo's counter-based index-draw spec.

Follows (file:line relative to /root/reference/serl_launcher/serl_launcher):
  * data/replay_buffer.py:41-75            ring storage, cursor/size bookkeeping
  * data/memory_efficient_replay_buffer.py:13-51   one frame per slot, validity mask
  * data/memory_efficient_replay_buffer.py:53-89   insert (episode-start fillers, wrap re-insert)
  * data/memory_efficient_replay_buffer.py:91-164  sample (window [idx-T, idx] -> packed frames)
  * data/dataset.py:79-102                 per-key gather at indx

PARITY PIN STATUS: pinned.  tests/golden/replay_*.npz were produced by running the REAL
reference classes (imported from /root/reference with gym/jax/flax import stubs, see
tests/golden/make_replay_golden.py) on scripted insert streams and scripted index streams;
tests/test_oracle_replay.py checks this restatement against them slot-for-slot.

Index draws: the reference draws from an UNSEEDED numpy Generator with sequential,
data-dependent redraws (memory_efficient_replay_buffer.py:111-122; dataset.py:55-69) and
`sample(indx=...)` raises NotImplementedError (:123-124), so no bit-exact reference stream
exists.  `draw_indices` below is THIS REPO's specification (Philox4x32-10 counter RNG +
Lemire's unbiased bounded integer + bounded redraw on invalid slots), implemented identically
by the CUDA sampler (serl_b200/csrc/sampler.cu).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline/reference legs may import this.
"""
from __future__ import annotations

from typing import Dict, Sequence

import numpy as np

U32 = np.uint32
U64 = np.uint64

# ----------------------------------------------------------------------------------------
# Philox4x32-10 (Salmon et al. 2011), scalar/array form
# ----------------------------------------------------------------------------------------
_PH_M0 = U64(0xD2511F53)
_PH_M1 = U64(0xCD9E8D57)
_PH_W0 = 0x9E3779B9
_PH_W1 = 0xBB67AE85


def philox4x32(ctr, key):
    """ctr: 4 uint32 arrays (broadcastable), key: 2 uint32 scalars -> 4 uint32 arrays."""
    c0, c1, c2, c3 = [np.asarray(c, dtype=U32) for c in ctr]
    c0, c1, c2, c3 = np.broadcast_arrays(c0, c1, c2, c3)
    k0 = int(key[0]) & 0xFFFFFFFF
    k1 = int(key[1]) & 0xFFFFFFFF
    for _ in range(10):
        p0 = c0.astype(U64) * _PH_M0
        p1 = c2.astype(U64) * _PH_M1
        hi0 = (p0 >> U64(32)).astype(U32)
        lo0 = (p0 & U64(0xFFFFFFFF)).astype(U32)
        hi1 = (p1 >> U64(32)).astype(U32)
        lo1 = (p1 & U64(0xFFFFFFFF)).astype(U32)
        n0 = hi1 ^ c1 ^ U32(k0)
        n1 = lo1
        n2 = hi0 ^ c3 ^ U32(k1)
        n3 = lo0
        c0, c1, c2, c3 = n0, n1, n2, n3
        k0 = (k0 + _PH_W0) & 0xFFFFFFFF
        k1 = (k1 + _PH_W1) & 0xFFFFFFFF
    return c0, c1, c2, c3


MAX_DRAW_ATTEMPTS = 64


def draw_indices(seed: int, step: int, batch: int, size: int, valid: np.ndarray,
                 lane_offset: int = 0) -> np.ndarray:
    """Repo spec for replay index draws (see module docstring).

    For lane i in [0, batch): attempt a = 0,1,...:
        x    = philox4x32(ctr=(lane_offset+i, a, step_lo, step_hi), key=(seed_lo, seed_hi))[0]
        m    = x * size (64-bit);  lo = m mod 2^32
        reject (Lemire) if lo < (2^32 - size) mod size
        idx  = m >> 32;  reject if not valid[idx]
    first accepted idx wins; after MAX_DRAW_ATTEMPTS rejections the lane yields -1
    (the product raises on that).  Returns int32 (batch,).
    """
    assert 0 < size <= valid.shape[0]
    out = np.full(batch, -1, dtype=np.int32)
    key = (seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    lanes = (np.arange(batch, dtype=np.int64) + lane_offset).astype(U32)
    pending = np.ones(batch, dtype=bool)
    thresh = ((1 << 32) - size) % size
    for a in range(MAX_DRAW_ATTEMPTS):
        if not pending.any():
            break
        x = philox4x32((lanes, U32(a), U32(step & 0xFFFFFFFF), U32((step >> 32) & 0xFFFFFFFF)), key)[0]
        m = x.astype(U64) * U64(size)
        lo = (m & U64(0xFFFFFFFF)).astype(np.int64)
        idx = (m >> U64(32)).astype(np.int64)
        ok = pending & (lo >= thresh) & valid[np.minimum(idx, size - 1)]
        out[ok] = idx[ok]
        pending &= ~ok
    return out


# ----------------------------------------------------------------------------------------
# Frame-dedup ring (restatement of MemoryEfficientReplayBuffer)
# ----------------------------------------------------------------------------------------
class OracleFrameRing:
    """One camera frame per slot; obs/next_obs pixel windows are reconstructed at sample time.

    Transition dict layout (what the actor sends, examples/async_drq_sim/async_drq_sim.py:145-152):
      observations:      {cam: (T,H,W,C) u8 ..., state: (T,S) f32}
      next_observations: same
      actions (A,), rewards (), masks (), dones ()
    """

    def __init__(self, capacity: int, image_keys: Sequence[str], frame_shape, num_stack: int,
                 state_dim: int, action_dim: int):
        self.capacity = int(capacity)
        self.image_keys = tuple(image_keys)
        self.T = int(num_stack)
        self.frames: Dict[str, np.ndarray] = {
            k: np.zeros((capacity, *frame_shape), dtype=np.uint8) for k in self.image_keys}
        self.state = np.zeros((capacity, self.T, state_dim), np.float32)
        self.next_state = np.zeros((capacity, self.T, state_dim), np.float32)
        self.actions = np.zeros((capacity, action_dim), np.float32)
        self.rewards = np.zeros((capacity,), np.float32)
        self.masks = np.zeros((capacity,), np.float32)
        self.dones = np.zeros((capacity,), bool)
        self.valid = np.zeros((capacity,), bool)      # reference: _is_correct_index
        self.size = 0                                   # reference: _size
        self.cursor = 0                                 # reference: _insert_index
        self.episode_start = True                       # reference: _first

    def __len__(self):
        return self.size

    # -- raw slot write: replay_buffer.py:71-75 ---------------------------------------------
    def _write_slot(self, frame_by_cam, state, next_state, action, reward, mask, done):
        i = self.cursor
        for k in self.image_keys:
            self.frames[k][i] = frame_by_cam[k]
        self.state[i] = state
        self.next_state[i] = next_state
        self.actions[i] = action
        self.rewards[i] = reward
        self.masks[i] = mask
        self.dones[i] = done
        self.cursor = (self.cursor + 1) % self.capacity
        self.size = min(self.size + 1, self.capacity)

    def insert(self, tr: dict):
        T = self.T
        # (:54-59) on wrap of a full buffer mid-episode, re-insert the last T slots at the front
        # as invalid copies so slot T's window [0..T] stays contiguous in time.
        if self.cursor == 0 and self.size == self.capacity and not self.episode_start:
            for src in range(self.size - T, self.size):
                self.valid[self.cursor] = False
                self._write_slot({k: self.frames[k][src].copy() for k in self.image_keys},
                                 self.state[src].copy(), self.next_state[src].copy(),
                                 self.actions[src].copy(), self.rewards[src], self.masks[src],
                                 self.dones[src])
        obs, nobs = tr["observations"], tr["next_observations"]
        st = np.asarray(obs["state"], np.float32).reshape(T, -1)
        nst = np.asarray(nobs["state"], np.float32).reshape(T, -1)
        a = np.asarray(tr["actions"], np.float32)
        r, m, d = np.float32(tr["rewards"]), np.float32(tr["masks"]), bool(tr["dones"])
        # (:71-77) episode start: T filler slots carrying the obs frames, marked invalid
        if self.episode_start:
            for t in range(T):
                self.valid[self.cursor] = False
                self._write_slot({k: np.asarray(obs[k])[t] for k in self.image_keys}, st, nst, a, r, m, d)
        # (:79-85) the transition slot stores the NEWEST next_obs frame and is valid
        self.episode_start = d
        self.valid[self.cursor] = True
        self._write_slot({k: np.asarray(nobs[k])[-1] for k in self.image_keys}, st, nst, a, r, m, d)
        # (:87-89) the T slots at the (new) cursor hold stale frames: their windows are broken
        for t in range(T):
            self.valid[(self.cursor + t) % self.size] = False

    # -- gather at explicit indices: (:126-164) with pack_obs_and_next_obs=True --------------
    def gather_packed(self, indx: np.ndarray) -> dict:
        indx = np.asarray(indx, dtype=np.int64)
        T = self.T
        out = {
            "observations": {"state": self.state[indx]},
            "next_observations": {"state": self.next_state[indx]},
            "actions": self.actions[indx],
            "rewards": self.rewards[indx],
            "masks": self.masks[indx],
            "dones": self.dones[indx],
        }
        # (:148-151) obs_pixels = sliding_window_view(frames, T + 1, axis=0)[indx - T]: the window axis has capacity - T entries and a
        # NEGATIVE index (a valid slot idx < T: the first transition of an episode whose filler frame landed on the last slots of the
        # ring) selects from its end, numpy-style - the reference returns slots capacity-T-1+(idx-T+1) .. , not the ring-wrapped
        # window.  Restated literally (pinned by tests/golden/replay_wrap_first.npz, generated by the real class).
        w0 = indx - T
        w0 = np.where(w0 < 0, w0 + (self.capacity - T), w0)
        win = w0[:, None] + np.arange(T + 1)[None, :]                  # slots of window w0
        for k in self.image_keys:
            out["observations"][k] = self.frames[k][win]              # (B, T+1, H, W, C)
        return out

    def sample(self, seed: int, step: int, batch: int, lane_offset: int = 0):
        idx = draw_indices(seed, step, batch, self.size, self.valid, lane_offset)
        if (idx < 0).any():
            raise RuntimeError("replay draw failed: no valid slot found within MAX_DRAW_ATTEMPTS")
        return idx, self.gather_packed(idx)


# ----------------------------------------------------------------------------------------
# Batch plumbing: utils/train_utils.py:16-31 (concat_batches), :44-66 (_unpack)
# ----------------------------------------------------------------------------------------
def concat_batches(first: dict, second: dict, axis: int = 0) -> dict:
    out = {}
    for k, v in first.items():
        out[k] = concat_batches(v, second[k], axis) if isinstance(v, dict) else np.concatenate((v, second[k]), axis)
    return out


def unpack(batch: dict) -> dict:
    """Packed (B,T+1,...) pixels -> obs = [:, :-1], next_obs = [:, 1:] where next_obs lacks the key."""
    obs = dict(batch["observations"])
    nobs = dict(batch["next_observations"])
    for k, v in batch["observations"].items():
        if k not in batch["next_observations"]:
            obs[k] = v[:, :-1]
            nobs[k] = v[:, 1:]
    out = dict(batch)
    out["observations"], out["next_observations"] = obs, nobs
    return out


# ----------------------------------------------------------------------------------------
# DrQ random shift: vision/data_augmentations.py:7-20 (edge-pad 4, dynamic_slice at (cy,cx))
# ----------------------------------------------------------------------------------------
def random_shift(frames: np.ndarray, offsets: np.ndarray, padding: int = 4) -> np.ndarray:
    """frames (N,H,W,C) u8, offsets (N,2) int [cy,cx] in [0, 2*padding] -> shifted frames.
    out[n,y,x] = frames[n, clip(y+cy-pad, 0, H-1), clip(x+cx-pad, 0, W-1)]."""
    N, H, W, _ = frames.shape
    ys = np.clip(np.arange(H)[None, :] + offsets[:, 0:1] - padding, 0, H - 1)
    xs = np.clip(np.arange(W)[None, :] + offsets[:, 1:2] - padding, 0, W - 1)
    return frames[np.arange(N)[:, None, None], ys[:, :, None], xs[:, None, :]]
