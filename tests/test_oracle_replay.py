"""Pins oracle/replay.py against fixtures produced by the REAL reference classes
(tests/golden/make_replay_golden.py) and checks the repo's index-draw spec."""
import glob
import os

import numpy as np
import pytest

from oracle import replay as R

GOLD = sorted(p for p in glob.glob(os.path.join(os.path.dirname(__file__), "golden", "replay_*.npz")) if not p.endswith("replay_wrap_first.npz"))
WRAP_FIRST = os.path.join(os.path.dirname(__file__), "golden", "replay_wrap_first.npz")


def _reference_index_protocol(stream, pos, B, size, valid):
    """memory_efficient_replay_buffer.py:111-122 driven by a scripted stream."""
    idx = np.array([stream[pos + i] % size for i in range(B)], dtype=np.int64)
    pos += B
    for i in range(B):
        while not valid[idx[i]]:
            idx[i] = stream[pos] % size
            pos += 1
    return idx, pos


@pytest.mark.parametrize("path", GOLD, ids=[os.path.basename(p) for p in GOLD])
def test_ring_matches_reference(path):
    g = np.load(path)
    cap, T, ncam, H, W, S, A, n_insert, B, n_batches = g["meta"].tolist()
    cams = [f"cam{i}" for i in range(ncam)]
    ring = R.OracleFrameRing(cap, cams, (H, W, 3), T, S, A)
    snaps = {int(g[f"snap{j}/n_inserted"]): f"snap{j}" for j in range(3)}
    for i in range(n_insert):
        tr = dict(
            observations={**{c: g[f"in_frames_{c}"][i] for c in cams}, "state": g["in_state"][i]},
            next_observations={**{c: g[f"in_nframes_{c}"][i] for c in cams}, "state": g["in_nstate"][i]},
            actions=g["in_actions"][i], rewards=g["in_rewards"][i], masks=g["in_masks"][i], dones=g["in_dones"][i])
        ring.insert(tr)
        if i + 1 in snaps:
            tag = snaps[i + 1]
            n = int(g[f"{tag}/size"])
            assert ring.size == n and ring.cursor == int(g[f"{tag}/cursor"])
            np.testing.assert_array_equal(ring.valid, g[f"{tag}/valid"])
            for name, arr in (("state", ring.state), ("next_state", ring.next_state), ("actions", ring.actions),
                              ("rewards", ring.rewards), ("masks", ring.masks), ("dones", ring.dones)):
                np.testing.assert_array_equal(arr[:n], g[f"{tag}/{name}"][:n], err_msg=name)
            for c in cams:
                np.testing.assert_array_equal(ring.frames[c][:n], g[f"{tag}/frames_{c}"][:n])
            stream = g[f"{tag}/stream"]
            pos = 0
            for b in range(n_batches):
                assert pos == int(g[f"{tag}_b{b}/pos0"])
                idx, pos = _reference_index_protocol(stream, pos, B, ring.size, ring.valid)
                assert pos == int(g[f"{tag}_b{b}/pos1"])
                out = ring.gather_packed(idx)
                for c in cams:
                    np.testing.assert_array_equal(out["observations"][c], g[f"{tag}_b{b}/pix_{c}"])
                    assert out["observations"][c].shape == (B, T + 1, H, W, 3)
                np.testing.assert_array_equal(out["observations"]["state"], g[f"{tag}_b{b}/state"])
                np.testing.assert_array_equal(out["next_observations"]["state"], g[f"{tag}_b{b}/next_state"])
                for k in ("actions", "rewards", "masks", "dones"):
                    np.testing.assert_array_equal(out[k], g[f"{tag}_b{b}/{k}"])


def test_valid_slot_below_T_reads_numpys_negative_window_like_the_reference():
    """An episode whose filler frame lands on the LAST slot puts its first (valid) transition on slot 0; the reference gathers
    sliding_window_view(frames)[idx - T], so idx = 0 returns window -1 = slots capacity-2, capacity-1 (fixture from the real class)."""
    g = np.load(WRAP_FIRST)
    cap, T, ncam, H, W, S, A, n_insert, B, _ = g["meta"].tolist()
    ring = R.OracleFrameRing(cap, ["cam0"], (H, W, 3), T, S, A)
    for i in range(n_insert):
        ring.insert(dict(observations={"cam0": g["in_frames_cam0"][i], "state": g["in_state"][i]},
                         next_observations={"cam0": g["in_nframes_cam0"][i], "state": g["in_nstate"][i]},
                         actions=g["in_actions"][i], rewards=g["in_rewards"][i], masks=g["in_masks"][i], dones=g["in_dones"][i]))
    np.testing.assert_array_equal(ring.valid, g["valid"])
    assert ring.valid[0] and ring.size == int(g["size"]) and ring.cursor == int(g["cursor"])
    np.testing.assert_array_equal(ring.frames["cam0"], g["frames_cam0"])
    idx = g["stream"].astype(np.int64) % ring.size
    assert ring.valid[idx].all() and (idx < T).any()
    out = ring.gather_packed(idx)
    np.testing.assert_array_equal(out["observations"]["cam0"], g["pix_cam0"])
    np.testing.assert_array_equal(out["observations"]["cam0"][0], ring.frames["cam0"][[cap - 2, cap - 1]])
    np.testing.assert_array_equal(out["observations"]["state"], g["state"])
    np.testing.assert_array_equal(out["next_observations"]["state"], g["next_state"])


def test_philox4x32_10_kat():
    # Random123 kat_vectors, philox4x32-10
    def ph(c, k):
        return tuple(int(v) for v in R.philox4x32([np.uint32(x) for x in c], k))
    assert ph([0, 0, 0, 0], (0, 0)) == (0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8)
    assert ph([0xFFFFFFFF] * 4, (0xFFFFFFFF, 0xFFFFFFFF)) == (0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD)
    assert ph([0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344], (0xA4093822, 0x299F31D0)) == \
        (0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1)


def test_draw_indices_only_valid_and_deterministic():
    rng = np.random.default_rng(0)
    valid = rng.random(1000) > 0.3
    valid[:1] = False
    a = R.draw_indices(7, 3, 512, 1000, valid)
    b = R.draw_indices(7, 3, 512, 1000, valid)
    assert (a == b).all() and valid[a].all() and a.min() >= 0
    c = R.draw_indices(7, 4, 512, 1000, valid)
    assert not (a == c).all()
    # lanes are independent counters: a lane offset reproduces the tail of a bigger draw
    d = R.draw_indices(7, 3, 256, 1000, valid, lane_offset=256)
    assert (d == a[256:]).all()
    # partial fill: never index beyond `size`
    e = R.draw_indices(1, 0, 256, 10, valid)
    assert e.max() < 10


def test_draw_indices_uniform_over_valid():
    valid = np.ones(97, bool)
    valid[::5] = False
    idx = np.concatenate([R.draw_indices(11, s, 1024, 97, valid) for s in range(40)])
    counts = np.bincount(idx, minlength=97)
    assert (counts[~valid] == 0).all()
    exp = idx.size / valid.sum()
    assert counts[valid].min() > 0.75 * exp and counts[valid].max() < 1.25 * exp


def test_draw_indices_no_valid_slot_flags_failure():
    assert (R.draw_indices(0, 0, 4, 8, np.zeros(8, bool)) == -1).all()


def test_random_shift_matches_pad_and_slice():
    rng = np.random.default_rng(1)
    fr = rng.integers(0, 256, (5, 9, 7, 3), dtype=np.uint8)
    off = np.array([[0, 0], [8, 8], [4, 4], [0, 8], [3, 6]], dtype=np.int32)
    out = R.random_shift(fr, off)
    for n in range(5):
        padded = np.pad(fr[n], ((4, 4), (4, 4), (0, 0)), mode="edge")   # data_augmentations.py:10-19
        cy, cx = off[n]
        np.testing.assert_array_equal(out[n], padded[cy:cy + 9, cx:cx + 7])
    np.testing.assert_array_equal(out[2], fr[2])     # centre offset is the identity


def test_concat_and_unpack():
    b1 = {"observations": {"cam": np.zeros((2, 2, 4, 4, 3), np.uint8), "state": np.zeros((2, 1, 3))},
          "next_observations": {"state": np.ones((2, 1, 3))}, "rewards": np.zeros(2)}
    b2 = {"observations": {"cam": np.ones((3, 2, 4, 4, 3), np.uint8), "state": np.ones((3, 1, 3))},
          "next_observations": {"state": np.ones((3, 1, 3))}, "rewards": np.ones(3)}
    c = R.concat_batches(b1, b2, axis=0)
    assert c["observations"]["cam"].shape == (5, 2, 4, 4, 3) and c["rewards"].tolist() == [0, 0, 1, 1, 1]
    u = R.unpack(c)
    assert u["observations"]["cam"].shape == (5, 1, 4, 4, 3)
    assert u["next_observations"]["cam"].shape == (5, 1, 4, 4, 3)
