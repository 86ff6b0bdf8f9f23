"""ORACLE (test infrastructure, not product): NumPy restatement of the JAX PRNG
primitives the DrQ/SAC path consumes.

The arithmetic lives in a third-party dependency that is NOT under /root/reference:
`jax` (README.md:56 pins jax[cuda12_pip]==0.4.35; CPU install unpinned).  This file
restates the published threefry2x32 algorithm and JAX's (non-"partitionable", the
default for jax<0.5) key-derivation layout.  Reference call sites it serves:

  * jax.random.split   - agents/continuous/sac.py:137,152,197,224,288;
                         agents/continuous/drq.py:279,308; common/common.py:199;
                         vision/data_augmentations.py:28
  * jax.random.randint - sac.py:153-158 (ensemble subsample),
                         vision/data_augmentations.py:8 (crop offsets)
  * jax.random.normal  - distrax MultivariateNormalDiag sampling (sac.py:128,201)
  * jax.random.bernoulli - flax.linen.Dropout (vision/resnet_v1.py:352)

PARITY PIN STATUS: the reference holds no golden vectors for these;
jax is not installable here.  The restatement is pinned against
  (a) the Random123 threefry2x32-20 known-answer vectors (also asserted by JAX's own
      random_test.py::testThreefry2x32), and
  (b) key/normal values printed in JAX's public documentation for PRNGKey(0)
      (split -> [4146024105 967050713], [2718843009 1272950319]; normal(key0,(1,)) ->
      -0.20584226; normal(subkey) -> -1.2515389),
see tests/test_oracle_prng.py.  Everything beyond that is "parity unpinned".

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline/reference legs may import this.
"""
from __future__ import annotations

import numpy as np

U32 = np.uint32
_ROT = ((13, 15, 26, 6), (17, 29, 16, 24))


def _rotl(x, d):
    return ((x << U32(d)) | (x >> U32(32 - d))).astype(U32)


def threefry2x32(key, x0, x1):
    """Threefry-2x32, 20 rounds.  key: (2,) uint32; x0, x1: uint32 arrays (same shape).
    Returns (y0, y1)."""
    with np.errstate(over="ignore"):
        k0 = U32(key[0])
        k1 = U32(key[1])
        ks = (k0, k1, U32(k0 ^ k1 ^ U32(0x1BD11BDA)))
        x0 = (np.asarray(x0, dtype=U32) + ks[0]).astype(U32)
        x1 = (np.asarray(x1, dtype=U32) + ks[1]).astype(U32)
        for r in range(5):
            for d in _ROT[r % 2]:
                x0 = (x0 + x1).astype(U32)
                x1 = _rotl(x1, d)
                x1 = (x1 ^ x0).astype(U32)
            x0 = (x0 + ks[(r + 1) % 3]).astype(U32)
            x1 = (x1 + ks[(r + 2) % 3] + U32(r + 1)).astype(U32)
    return x0, x1


def prng_key(seed: int) -> np.ndarray:
    """jax.random.PRNGKey(seed) for the threefry impl: [hi32, lo32]."""
    seed = int(seed) & 0xFFFFFFFFFFFFFFFF
    return np.array([(seed >> 32) & 0xFFFFFFFF, seed & 0xFFFFFFFF], dtype=U32)


def _threefry_counts(key, counts):
    """jax._src.prng.threefry_2x32(keypair, count): split the flat counter array in two
    halves (padding one zero if odd), hash pairwise, concatenate."""
    counts = np.asarray(counts, dtype=U32).ravel()
    n = counts.size
    odd = n % 2
    if odd:
        counts = np.concatenate([counts, np.zeros(1, U32)])
    h = counts.size // 2
    y0, y1 = threefry2x32(key, counts[:h], counts[h:])
    out = np.concatenate([y0, y1])
    return out[:-1] if odd else out


def split(key, num: int = 2) -> np.ndarray:
    """jax.random.split(key, num) -> (num, 2) uint32 (original, non-partitionable layout)."""
    return _threefry_counts(key, np.arange(2 * num, dtype=U32)).reshape(num, 2)


def fold_in(key, data: int) -> np.ndarray:
    """jax.random.fold_in(key, data) = threefry(key, PRNGKey(data))."""
    d = prng_key(data)
    y0, y1 = threefry2x32(key, d[0:1], d[1:2])
    return np.array([y0[0], y1[0]], dtype=U32)


def random_bits(key, shape) -> np.ndarray:
    """32-bit random bits of `shape` (jax _threefry_random_bits_original, bit_width=32)."""
    size = int(np.prod(shape, dtype=np.int64)) if len(shape) else 1
    return _threefry_counts(key, np.arange(size, dtype=U32)).reshape(shape)


def randint(key, shape, minval: int, maxval: int) -> np.ndarray:
    """jax.random.randint(key, shape, minval, maxval) for int32, maxval > minval."""
    k1, k2 = split(key, 2)
    hb = random_bits(k1, shape).astype(np.uint64)
    lb = random_bits(k2, shape).astype(np.uint64)
    span = np.uint64(maxval - minval)
    mult = np.uint64(2**16) % span
    mult = (mult * mult) % span
    # uint32 arithmetic in JAX; (hb%span)*mult + lb%span cannot overflow 32 bits for the
    # spans used on this path (9, 10) - keep an assertion for anything larger.
    assert int(span) * int(span) < 2**32
    off = ((hb % span) * mult + (lb % span)) % span
    return (np.int64(minval) + off.astype(np.int64)).astype(np.int32)


def uniform01(key, shape) -> np.ndarray:
    """jax.random.uniform(key, shape, float32) in [0, 1)."""
    bits = random_bits(key, shape)
    fb = ((bits >> U32(9)) | U32(0x3F800000)).astype(U32)
    return (fb.view(np.float32) - np.float32(1.0)).astype(np.float32)


def erfinv_f32(x: np.ndarray) -> np.ndarray:
    """Single-precision erf^-1 (Giles 2010 polynomial, the form XLA lowers f32 erf_inv to).
    Evaluated in float32 like the device kernel."""
    x = np.asarray(x, dtype=np.float32)
    w = (-np.log1p((-x * x).astype(np.float32))).astype(np.float32)
    lt = w < np.float32(5.0)
    w1 = (w - np.float32(2.5)).astype(np.float32)
    w2 = (np.sqrt(np.maximum(w, 0)).astype(np.float32) - np.float32(3.0)).astype(np.float32)

    def poly(cs, ww):
        p = np.full_like(ww, np.float32(cs[0]))
        for c in cs[1:]:
            p = (np.float32(c) + p * ww).astype(np.float32)
        return p

    c1 = (2.81022636e-08, 3.43273939e-07, -3.5233877e-06, -4.39150654e-06, 0.00021858087,
          -0.00125372503, -0.00417768164, 0.246640727, 1.50140941)
    c2 = (-0.000200214257, 0.000100950558, 0.00134934322, -0.00367342844, 0.00573950773,
          -0.0076224613, 0.00943887047, 1.00167406, 2.83297682)
    p = np.where(lt, poly(c1, w1), poly(c2, w2)).astype(np.float32)
    return (p * x).astype(np.float32)


def normal(key, shape) -> np.ndarray:
    """jax.random.normal(key, shape, float32): sqrt(2) * erfinv(uniform(-1+ulp, 1))."""
    lo = np.nextafter(np.float32(-1.0), np.float32(0.0))
    hi = np.float32(1.0)
    f = uniform01(key, shape)
    u = np.maximum(lo, (f * np.float32(hi - lo) + lo).astype(np.float32)).astype(np.float32)
    return (np.float32(np.sqrt(2.0)) * erfinv_f32(u)).astype(np.float32)


def bernoulli(key, p: float, shape) -> np.ndarray:
    """jax.random.bernoulli(key, p, shape) = uniform(key, shape) < p."""
    return uniform01(key, shape) < np.float32(p)


# ----------------------------------------------------------------------------------------
# Derived draws used by the DrQ step
# ----------------------------------------------------------------------------------------

def crop_offsets(key, n_frames: int, padding: int = 4) -> np.ndarray:
    """vision/data_augmentations.py:22-36 + :7-9: rngs = split(key, n_frames); per frame
    crop_from = randint(rng_i, (2,), 0, 2*padding+1) -> (cy, cx).  Returns (n_frames, 2) int32."""
    keys = split(key, n_frames)
    return np.stack([randint(k, (2,), 0, 2 * padding + 1) for k in keys]).astype(np.int32)
