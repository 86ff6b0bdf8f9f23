"""Flat parameter layout and initialisers (serl_b200/params.py) against the figures of SURVEY.md §8d / App. D."""
import math

import numpy as np

from serl_b200 import params as P


def _count(spec):
    return sum(l.size for l in spec)


def test_trainable_counts_match_survey():
    one = P.trainable_spec(("front",), 7, 4, 10, True)
    two = P.trainable_spec(("front", "wrist"), 7, 4, 10, True)
    assert abs(_count(one) - 2.77e6) < 0.01e6 and abs(_count(two) - 4.60e6) < 0.01e6
    crit = sum(l.size for l in one if l.group == 0)
    assert abs(crit - 2.62e6) < 0.01e6                                                       # critic-step all-reduce payload, 10.5 MB
    trunk = sum(int(np.prod(s)) for _, s in P.trunk_spec(3))
    assert abs(trunk - 4.90e6) < 0.01e6                                                      # frozen ResNet-10 per camera


def test_flat_layout_is_group_major_and_16_byte_aligned():
    spec = P.trainable_spec(("a", "b"), 7, 4, 10, True)
    groups = [l.group for l in spec]
    assert groups == sorted(groups) and set(groups) == {0, 1, 2}          # optimizer 0 (critic tx), 1 (actor), 2 (temperature)
    end = 0
    for l in spec:
        assert l.offset % 4 == 0 and l.offset >= end                     # 4 floats = 16 bytes; leaves do not overlap
        end = l.offset + l.size
    assert spec[-1].path == "modules_temperature/lagrange" and spec[-1].shape == ()


def test_pixel_and_state_critic_heads_differ_like_the_reference():
    pix = {l.path: l.shape for l in P.trainable_spec(("a",), 7, 4, 10, True)}
    st = {l.path: l.shape for l in P.trainable_spec((), 10, 4, 2, False)}
    assert pix["modules_critic/Dense_0/kernel"] == (256, 1)              # one shared value head (drq.py:201-207)
    assert st["modules_critic/Dense_0/kernel"] == (2, 256, 1)            # whole critic ensembled (sac.py:523-524)
    assert pix["modules_critic/network/Dense_0/kernel"] == (10, 256 + 64 + 4, 256)
    assert st["modules_critic/network/Dense_0/kernel"] == (2, 10 + 4, 256)


def test_initialiser_distributions():
    rng = np.random.default_rng(0)
    w = P.xavier_uniform(rng, (320, 256))
    lim = math.sqrt(6.0 / (320 + 256))
    assert w.dtype == np.float32 and abs(w).max() <= lim and abs(w.std() - lim / math.sqrt(3)) < 0.02 * lim
    k = P.kaiming_normal(rng, (3, 3, 64, 128))                           # fan_in = 3*3*64; truncated at 2 sigma, variance restored
    assert abs(k.std() - math.sqrt(2.0 / 576)) < 0.02 * math.sqrt(2.0 / 576)
    assert abs(k).max() <= 2.0 * math.sqrt(2.0 / 576) / 0.87962566103423978 + 1e-6
    l = P.lecun_normal(rng, (4096, 256))
    assert abs(l.std() - math.sqrt(1.0 / 4096)) < 0.02 * math.sqrt(1.0 / 4096)
    t = P.init_trainable(rng, P.trainable_spec(("a",), 7, 4, 10, True), temperature_init=1.0)
    assert abs(float(np.log1p(np.exp(t["modules_temperature/lagrange"]))) - 1.0) < 1e-6      # softplus(lagrange) = temperature_init
    members = t["modules_critic/network/Dense_1/kernel"]
    assert members.shape == (10, 256, 256) and not np.allclose(members[0], members[1])      # vmapped members initialised independently
