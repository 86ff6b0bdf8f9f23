"""CPU: the C-ABI library builds, loads and exports every symbol include/serl_b200.h declares; host mirrors
of the device PRNG agree with the oracle (no GPU compute calls)."""
import os
import re

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _lib():
    import __graft_entry__ as G
    G.build()
    from serl_b200 import _lib as L
    return L, L.load()


def test_header_symbols_exported_and_bound():
    L, lib = _lib()
    hdr = open(os.path.join(ROOT, "include", "serl_b200.h")).read()
    declared = set(re.findall(r"\b(serl_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no declarations parsed"
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in serl_b200.h but not exported"
    assert declared == set(L.EXPORTS), (declared ^ set(L.EXPORTS))
    assert lib.serl_version() == L.ABI_VERSION


def test_struct_layouts_match_header_sizes():
    import ctypes as C
    L, _ = _lib()
    # spot-check ABI sizes against the C layout rules the header implies (LP64)
    assert C.sizeof(L.ReplayView) == 8 * 4 + 8 * 7 + 4 * 9 + 4      # 4 cam ptrs + 7 ptrs + 9 int32 (+4 pad)
    assert C.sizeof(L.GemmDesc) % 8 == 0 and C.sizeof(L.AdamDesc) % 8 == 0


def test_missing_library_fails_loudly(tmp_path, monkeypatch):
    from serl_b200 import _lib as L
    monkeypatch.setattr(L, "_lib", None)
    monkeypatch.setattr(L, "LIB_PATH", str(tmp_path / "nope.so"))
    try:
        L.load()
    except L.SerlError as e:
        assert "no fallback" in str(e)
    else:
        raise AssertionError("loading a missing library must raise")


def test_host_mirrors_match_oracle():
    from oracle import jax_prng as P
    from oracle import replay as R
    L, _ = _lib()
    key = P.prng_key(1234)
    for n in (1, 2, 3, 7, 256, 512):
        out = np.zeros((n, 2), np.uint32)
        L.call("serl_host_threefry_split", key.ctypes.data, n, out.ctypes.data)
        np.testing.assert_array_equal(out, P.split(key, n))
        off = np.zeros((n, 2), np.int32)
        L.call("serl_host_crop_offsets", key.ctypes.data, n, 4, off.ctypes.data)
        np.testing.assert_array_equal(off, P.crop_offsets(key, n))
        bits = np.zeros(n, np.uint32)
        L.call("serl_host_random_bits", key.ctypes.data, n, bits.ctypes.data)
        np.testing.assert_array_equal(bits, P.random_bits(key, (n,)))
    valid = (np.random.default_rng(0).random(1000) > 0.3).astype(np.uint8)
    out = np.zeros(512, np.int32)
    L.call("serl_host_draw_indices", 7, 3, 0, 512, 1000, valid.ctypes.data, out.ctypes.data)
    np.testing.assert_array_equal(out, R.draw_indices(7, 3, 512, 1000, valid.astype(bool)))
    out2 = np.zeros(100, np.int32)
    L.call("serl_host_draw_indices", (5 << 32) | 9, (1 << 33) + 5, 17, 100, 777, valid.ctypes.data, out2.ctypes.data)
    np.testing.assert_array_equal(out2, R.draw_indices((5 << 32) | 9, (1 << 33) + 5, 100, 777, valid.astype(bool), lane_offset=17))


def test_host_key_schedule_matches_reference_split_order():
    from oracle import jax_prng as P
    L, _ = _lib()
    r = P.prng_key(42).copy()
    keys = np.zeros(2 * L.NUM_KEYS, np.uint32)
    L.call("serl_host_rng_schedule", r.ctypes.data, keys.ctypes.data, 1, 1)
    r1, ko, kn = P.split(P.prng_key(42), 3)                 # drq.py:307-308
    _, ka, kc, kt = P.split(r1, 4)                          # common.py:198-200 (actor, critic, temperature)
    c1, kna = P.split(kc)                                   # sac.py:137
    _, ksub = P.split(c1)                                   # sac.py:152
    _, kp, ks, _ = P.split(ka, 4)                           # sac.py:197
    _, ktn = P.split(kt)                                    # sac.py:224
    got = keys.reshape(-1, 2)
    for slot, ref in ((L.KEY_CROP_OBS, ko), (L.KEY_CROP_NEXT, kn), (L.KEY_CRITIC_NEXT, kna), (L.KEY_CRITIC_SUBSAMPLE, ksub),
                      (L.KEY_ACTOR_DROPOUT, kp), (L.KEY_ACTOR_SAMPLE, ks), (L.KEY_TEMP_NEXT, ktn)):
        np.testing.assert_array_equal(got[slot], ref)
    np.testing.assert_array_equal(r, P.split(r1)[0])        # sac.py:288
