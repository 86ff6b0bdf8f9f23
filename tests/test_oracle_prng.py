"""Pins oracle/jax_prng.py (restated JAX threefry PRNG) against known-answer vectors."""
import numpy as np

from oracle import jax_prng as P


def _tf(key, ctr):
    y0, y1 = P.threefry2x32(np.array(key, np.uint32), np.array([ctr[0]], np.uint32), np.array([ctr[1]], np.uint32))
    return int(y0[0]), int(y1[0])


def test_threefry2x32_random123_kat():
    # Random123 kat_vectors (threefry2x32, 20 rounds); the same three vectors are asserted by
    # jax/tests/random_test.py::testThreefry2x32.
    assert _tf([0, 0], [0, 0]) == (0x6B200159, 0x99BA4EFE)
    assert _tf([0xFFFFFFFF] * 2, [0xFFFFFFFF] * 2) == (0x1CB996FC, 0xBB002BE7)
    assert _tf([0x13198A2E, 0x03707344], [0x243F6A88, 0x85A308D3]) == (0xC4923A9C, 0x483DF7A0)


def test_split_and_normal_match_jax_docs():
    # Values printed in JAX's public docs ("Sharp bits" / PRNG tutorial) for PRNGKey(0).
    k = P.prng_key(0)
    assert k.tolist() == [0, 0]
    new, sub = P.split(k)
    assert new.tolist() == [4146024105, 967050713]
    assert sub.tolist() == [2718843009, 1272950319]
    np.testing.assert_allclose(P.normal(k, (1,)), [-0.20584226], rtol=0, atol=1e-7)
    np.testing.assert_allclose(P.normal(sub, (1,)), [-1.2515389], rtol=0, atol=2e-7)
    new2, sub2 = P.split(new)
    assert new2.tolist() == [2384771982, 3928867769]
    assert sub2.tolist() == [1278412471, 2182328957]
    np.testing.assert_allclose(P.normal(sub2, (1,)), [-0.58665055], rtol=0, atol=1e-7)


def test_prng_key_packs_hi_lo():
    assert P.prng_key(42).tolist() == [0, 42]
    assert P.prng_key((7 << 32) | 5).tolist() == [7, 5]


def test_randint_range_and_uniformity():
    k = P.prng_key(123)
    v = P.randint(k, (9000,), 0, 9)
    assert v.min() == 0 and v.max() == 8 and v.dtype == np.int32
    counts = np.bincount(v, minlength=9)
    assert counts.min() > 850 and counts.max() < 1150
    w = P.randint(k, (2,), 0, 10)
    assert ((0 <= w) & (w < 10)).all()


def test_random_bits_odd_size_padding():
    k = P.prng_key(9)
    b5 = P.random_bits(k, (5,))
    # odd sizes pad the counter array with one zero before halving
    y0, y1 = P.threefry2x32(k, np.array([0, 1, 2], np.uint32), np.array([3, 4, 0], np.uint32))
    assert b5.tolist() == np.concatenate([y0, y1])[:5].tolist()


def test_crop_offsets_shape_and_determinism():
    k = P.prng_key(5)
    a = P.crop_offsets(k, 6)
    b = P.crop_offsets(k, 6)
    assert a.shape == (6, 2) and (a == b).all() and a.min() >= 0 and a.max() <= 8
    # frame i's key is row i of split(key, n): changing n changes every key (JAX layout)
    assert not (P.crop_offsets(k, 7)[:6] == a).all()


def test_bernoulli_rate():
    m = P.bernoulli(P.prng_key(3), 0.9, (20000,))
    assert 0.89 < m.mean() < 0.91
