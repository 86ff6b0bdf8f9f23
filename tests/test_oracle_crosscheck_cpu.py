"""The network arithmetic of oracle/drq.py restates third-party definitions (flax LayerNorm / GroupNorm / Conv SAME padding,
optax Adam, distrax tanh-Gaussian) that cannot be imported here.  These tests check each building block against an
INDEPENDENT implementation of the same published definition that ships with torch (or a literal loop), so a slip in the
restatement cannot hide behind the GPU parity tests, which compare the kernels with this same oracle."""
import math

import numpy as np
import torch
import torch.nn.functional as F

from oracle import drq as O

torch.manual_seed(0)


def test_layer_norm_matches_torch_layer_norm():
    x = torch.randn(7, 5, 256, dtype=torch.float64) * 3 + 0.7
    s, b = torch.randn(256, dtype=torch.float64), torch.randn(256, dtype=torch.float64)
    ref = F.layer_norm(x, (256,), s, b, eps=1e-6)
    assert (O.layer_norm(x, s, b) - ref).abs().max() < 1e-11


def test_group_norm_matches_torch_group_norm():
    x = torch.randn(3, 8, 8, 64, dtype=torch.float64) * 2 - 0.3
    s, b = torch.randn(64, dtype=torch.float64), torch.randn(64, dtype=torch.float64)
    ref = F.group_norm(x.permute(0, 3, 1, 2), 4, s, b, eps=1e-5).permute(0, 2, 3, 1)
    assert (O.group_norm_nhwc(x, s, b) - ref).abs().max() < 1e-11


def test_same_padding_is_xla_same():
    # XLA SAME: out = ceil(size / stride), total = max((out-1)*stride + k - size, 0), low = total // 2 (the extra cell goes high)
    assert O.same_pads(128, 7, 2) == (2, 3) and O.same_pads(64, 3, 2) == (0, 1) and O.same_pads(32, 3, 1) == (1, 1)
    assert O.same_pads(32, 1, 2) == (0, 0) and O.same_pads(5, 3, 2) == (1, 1)


def test_conv_and_maxpool_against_literal_loops():
    rng = np.random.default_rng(1)
    x = rng.standard_normal((2, 6, 6, 3))
    w = rng.standard_normal((3, 3, 3, 4))
    lo, hi = O.same_pads(6, 3, 2)
    got = O.conv_nhwc(torch.as_tensor(x), torch.as_tensor(w), 2, lo, hi).numpy()
    xp = np.pad(x, ((0, 0), (lo, hi), (lo, hi), (0, 0)))
    ref = np.zeros((2, 3, 3, 4))
    for i in range(3):
        for j in range(3):
            ref[:, i, j] = np.einsum("nhwc,hwco->no", xp[:, 2 * i:2 * i + 3, 2 * j:2 * j + 3], w)
    np.testing.assert_allclose(got, ref, atol=1e-12)
    mp = O.max_pool_3x3_s2_same(torch.as_tensor(x)).numpy()
    xm = np.pad(x, ((0, 0), (lo, hi), (lo, hi), (0, 0)), constant_values=-np.inf)
    refm = np.stack([np.stack([xm[:, 2 * i:2 * i + 3, 2 * j:2 * j + 3].max(axis=(1, 2)) for j in range(3)], 1) for i in range(3)], 1)
    np.testing.assert_array_equal(mp, refm)


def test_adam_update_matches_torch_adam():
    # optax.adam: m_hat / (sqrt(v_hat) + eps) with bias correction - the formulation torch.optim.Adam implements as well
    p0 = torch.randn(50, dtype=torch.float64)
    p_ref = p0.clone().requires_grad_(True)
    opt_ref = torch.optim.Adam([p_ref], lr=3e-4, betas=(0.9, 0.999), eps=1e-8)
    p = p0.clone()
    opt = {"count": 0, "mu": {"w": torch.zeros_like(p)}, "nu": {"w": torch.zeros_like(p)}}
    for step in range(6):
        g = torch.randn(50, dtype=torch.float64) * (0.0 if step == 3 else 1.0)     # one zero-gradient step: momentum keeps moving p
        p = p + O.adam_tx_update({"w": g}, opt, 3e-4)["w"]
        p_ref.grad = g.clone()
        opt_ref.step()
        assert (p - p_ref.detach()).abs().max() < 1e-14
    assert opt["count"] == 6


def test_tanh_gaussian_logp_matches_torch_distributions():
    from torch.distributions import Independent, Normal, TransformedDistribution
    from torch.distributions.transforms import TanhTransform
    mu = torch.randn(9, 4, dtype=torch.float64)
    std = torch.rand(9, 4, dtype=torch.float64) + 0.1
    eps = torch.randn(9, 4, dtype=torch.float64)
    a, logp = O.tanh_normal_sample_logp(mu, std, eps)
    dist = TransformedDistribution(Independent(Normal(mu, std), 1), [TanhTransform(cache_size=1)])
    u = mu + std * eps
    ref = Independent(Normal(mu, std), 1).log_prob(u) - TanhTransform().log_abs_det_jacobian(u, torch.tanh(u)).sum(-1)
    assert (a - torch.tanh(u)).abs().max() == 0
    assert (logp - ref).abs().max() < 1e-12
    assert (dist.log_prob(a) - logp).abs().max() < 1e-6          # through atanh(a): looser, but an end-to-end check of the density


def test_lr_schedule_is_linear_warmup_then_constant():
    assert O.lr_schedule(0, 3e-4, 10) == 0.0 and abs(O.lr_schedule(5, 3e-4, 10) - 1.5e-4) < 1e-18
    assert O.lr_schedule(10, 3e-4, 10) == 3e-4 and O.lr_schedule(7, 3e-4, 0) == 3e-4
