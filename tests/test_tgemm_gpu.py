"""serl_tgemm_tf32 (csrc/tgemm.cu): single-pass TF32 GEMM with TMA-fed operands in every operand layout the heads use, and its
fused epilogues, against float64 references.  Tolerance = TF32 operand rounding (2^-11 relative per operand, random signs)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

TOL = 2e-3          # max |err| / max |ref|; measured ~3e-4 for K = 4096


def cu(x):
    return torch.as_tensor(x).cuda()


def rel_err(a, b):
    return float(np.max(np.abs(np.asarray(a, np.float64) - b)) / (np.max(np.abs(b)) + 1e-30))


@pytest.mark.parametrize("M,K,N,Z", [(256, 580, 256, 10), (37, 4096, 256, 1), (256, 256, 256, 3), (300, 576, 256, 1), (32, 580, 256, 10)])
def test_forward_x_w_layouts(M, K, N, Z):
    """X (M,K) row-major [K-major A] @ W (K,N) row-major [MN-major B] + bias; automatic k-split for the tall K."""
    from serl_b200 import ops
    rng = np.random.default_rng(5)
    x = rng.standard_normal((M, K)).astype(np.float32)                      # shared by the Z members (z stride 0)
    w = (rng.standard_normal((Z, K, N)) / np.sqrt(K)).astype(np.float32)
    b = rng.standard_normal((Z, N)).astype(np.float32)
    xd, wd, bd = cu(x), cu(w), cu(b)
    out = torch.full((Z, M, N), 7.0, device="cuda")
    err = torch.zeros(1, dtype=torch.int32, device="cuda")
    ws = ops.Workspace(64 << 20, "cuda")
    p = ops.tgemm_problem(xd.data_ptr(), wd.data_ptr(), sAm=K, sAk=1, sBk=N, sBn=1, Z=Z, sAz=0, sBz=K * N, C_=out.data_ptr(), sCz=M * N, ldc=N,
                          bias=bd.data_ptr(), sBiasZ=N)
    ops.tgemm(ws, [p], M, N, K, error=err)
    ref = np.einsum("mk,zkn->zmn", x.astype(np.float64), w.astype(np.float64)) + b[:, None, :]
    assert int(err.item()) == 0
    assert rel_err(out.cpu().numpy(), ref) < TOL


@pytest.mark.parametrize("M,K,N,Z", [(256, 256, 580, 10), (256, 256, 4096, 1), (40, 256, 256, 2)])
def test_input_gradient_layout_and_reduce_z(M, K, N, Z):
    """dX (M, N=fan_in) = dZ (M, K=256) [K-major A] @ W^T with W (fan_in, 256) row-major [K-major B]; sum over members."""
    from serl_b200 import ops
    rng = np.random.default_rng(6)
    dz = rng.standard_normal((Z, M, K)).astype(np.float32)
    w = (rng.standard_normal((Z, N, K)) / np.sqrt(K)).astype(np.float32)    # (fan_in, 256)
    dzd, wd = cu(dz), cu(w)
    ws = ops.Workspace(64 << 20, "cuda")
    ref = np.einsum("zmk,znk->zmn", dz.astype(np.float64), w.astype(np.float64))
    out = torch.zeros(Z, M, N, device="cuda")
    p = ops.tgemm_problem(dzd.data_ptr(), wd.data_ptr(), sAm=K, sAk=1, sBk=1, sBn=K, Z=Z, sAz=M * K, sBz=N * K, C_=out.data_ptr(), sCz=M * N, ldc=N)
    ops.tgemm(ws, [p], M, N, K)
    assert rel_err(out.cpu().numpy(), ref) < TOL
    red = torch.zeros(M, N, device="cuda")
    p = ops.tgemm_problem(dzd.data_ptr(), wd.data_ptr(), sAm=K, sAk=1, sBk=1, sBn=K, Z=Z, sAz=M * K, sBz=N * K, C_=red.data_ptr(), sCz=0, ldc=N)
    ops.tgemm(ws, [p], M, N, K, reduce_z=True)
    assert rel_err(red.cpu().numpy(), ref.sum(0)) < TOL


@pytest.mark.parametrize("R,FI,N,Z", [(256, 580, 256, 10), (256, 4096, 256, 1), (32, 256, 256, 10)])
def test_weight_gradient_layout(R, FI, N, Z):
    """dW (fan_in, 256) = X^T dZ: A = X (R, fan_in) row-major read MN-major, B = dZ (R, 256) row-major read MN-major; k = batch rows."""
    from serl_b200 import ops
    rng = np.random.default_rng(7)
    x = rng.standard_normal((R, FI)).astype(np.float32)
    dz = rng.standard_normal((Z, R, N)).astype(np.float32)
    xd, dzd = cu(x), cu(dz)
    ws = ops.Workspace(64 << 20, "cuda")
    dw = torch.zeros(Z, FI, N, device="cuda")
    p = ops.tgemm_problem(xd.data_ptr(), dzd.data_ptr(), sAm=1, sAk=FI, sBk=N, sBn=1, Z=Z, sAz=0, sBz=R * N, C_=dw.data_ptr(), sCz=FI * N, ldc=N)
    ops.tgemm(ws, [p], FI, N, R)
    ref = np.einsum("rk,zrn->zkn", x.astype(np.float64), dz.astype(np.float64))
    assert rel_err(dw.cpu().numpy(), ref) < TOL


def _ln_tanh(zv, scale, bias, eps=1e-6):
    mean = zv.mean(-1, keepdims=True)
    var = np.maximum((zv * zv).mean(-1, keepdims=True) - mean * mean, 0.0)
    rstd = 1.0 / np.sqrt(var + eps)
    xh = (zv - mean) * rstd
    return np.tanh(xh * scale + bias), xh, rstd[..., 0]


def test_layernorm_tanh_value_head_epilogue_multi_problem():
    """Two problems in one launch (different inputs and weights), Dense + bias + LayerNorm + tanh + value head, saves."""
    from serl_b200 import _lib as L
    from serl_b200 import ops
    rng = np.random.default_rng(8)
    M, K, N, Z = 200, 580, 256, 10
    probs, refs, outs = [], [], []
    keep = []
    for i in range(2):
        x = rng.standard_normal((M, K)).astype(np.float32)
        w = (rng.standard_normal((Z, K, N)) / np.sqrt(K)).astype(np.float32)
        b = (0.1 * rng.standard_normal((Z, N))).astype(np.float32)
        sc = (1 + 0.1 * rng.standard_normal((Z, N))).astype(np.float32)
        lb = (0.1 * rng.standard_normal((Z, N))).astype(np.float32)
        hw = (rng.standard_normal((N, 1)) / 16).astype(np.float32)
        hb = rng.standard_normal(1).astype(np.float32)
        t = [cu(v) for v in (x, w, b, sc, lb, hw, hb)]
        h = torch.zeros(Z, M, N, device="cuda"); xh = torch.zeros(Z, M, N, device="cuda"); rs = torch.zeros(Z, M, device="cuda")
        q = torch.zeros(Z, M, device="cuda")
        keep.append(t)
        probs.append(ops.tgemm_problem(t[0].data_ptr(), t[1].data_ptr(), sAm=K, sAk=1, sBk=N, sBn=1, Z=Z, sAz=0, sBz=K * N,
                                       C_=h.data_ptr(), sCz=M * N, ldc=N, bias=t[2].data_ptr(), sBiasZ=N, ln_scale=t[3].data_ptr(),
                                       ln_bias=t[4].data_ptr(), sLnZ=N, xhat=xh.data_ptr(), rstd=rs.data_ptr(), sXhatZ=M * N, sRstdZ=M,
                                       head_w=t[5].data_ptr(), head_b=t[6].data_ptr(), head_out=q.data_ptr(), sHeadOutZ=M, ld_head=1))
        zv = np.einsum("mk,zkn->zmn", x.astype(np.float64), w.astype(np.float64)) + b[:, None, :]
        hr, xr, rr = _ln_tanh(zv, sc[:, None, :].astype(np.float64), lb[:, None, :].astype(np.float64))
        refs.append((hr, xr, rr, hr @ hw.astype(np.float64)[:, 0] + hb[0]))
        outs.append((h, xh, rs, q))
    ops.tgemm(None, probs, M, N, K, epilogue=L.TGEMM_LN_TANH_HEAD, head_n=1)
    for (h, xh, rs, q), (hr, xr, rr, qr) in zip(outs, refs):
        assert rel_err(h.cpu().numpy(), hr) < TOL
        assert rel_err(xh.cpu().numpy(), xr) < 2 * TOL
        assert rel_err(rs.cpu().numpy(), rr) < TOL
        assert rel_err(q.cpu().numpy(), qr) < TOL


def test_policy_epilogue_matches_separate_kernels():
    """Dense + LayerNorm + tanh + mean / log-std heads + tanh-Gaussian sample in the epilogue == the separate kernels fed
    with the epilogue's own h (the head / sample arithmetic is fp32 in both)."""
    from serl_b200 import _lib as L
    from serl_b200 import ops
    rng = np.random.default_rng(9)
    M, K, N, A = 150, 256, 256, 4
    x = np.tanh(rng.standard_normal((M, K))).astype(np.float32)
    w = (rng.standard_normal((K, N)) / np.sqrt(K)).astype(np.float32)
    b = (0.1 * rng.standard_normal(N)).astype(np.float32)
    sc, lb = (1 + 0.1 * rng.standard_normal(N)).astype(np.float32), (0.1 * rng.standard_normal(N)).astype(np.float32)
    wm, wl = (rng.standard_normal((N, A)) / 16).astype(np.float32), (rng.standard_normal((N, A)) / 16).astype(np.float32)
    bm, bl = rng.standard_normal(A).astype(np.float32), rng.standard_normal(A).astype(np.float32)
    eps = rng.standard_normal((M, A)).astype(np.float32)
    t = {k: cu(v) for k, v in dict(x=x, w=w, b=b, sc=sc, lb=lb, wm=wm, wl=wl, bm=bm, bl=bl, eps=eps).items()}
    h = torch.zeros(M, N, device="cuda")
    mu, ls, u, sd = (torch.zeros(M, A, device="cuda") for _ in range(4))
    act = torch.zeros(M, A + 3, device="cuda"); logp = torch.zeros(M, device="cuda")
    p = ops.tgemm_problem(t["x"].data_ptr(), t["w"].data_ptr(), sAm=K, sAk=1, sBk=N, sBn=1, C_=h.data_ptr(), ldc=N, bias=t["b"].data_ptr(),
                          ln_scale=t["sc"].data_ptr(), ln_bias=t["lb"].data_ptr(), head_w=t["wm"].data_ptr(), head_b=t["bm"].data_ptr(),
                          head_out=mu.data_ptr(), head_w2=t["wl"].data_ptr(), head_b2=t["bl"].data_ptr(), head_out2=ls.data_ptr(),
                          noise=t["eps"].data_ptr(), act=act.data_ptr(), ld_act=A + 3, logp=logp.data_ptr(), u_out=u.data_ptr(), std_out=sd.data_ptr())
    ops.tgemm(None, [p], M, N, K, epilogue=L.TGEMM_LN_TANH_POLICY, head_n=A, std_min=1e-5, std_max=5.0)
    hn = h.cpu().numpy().astype(np.float64)
    zv = x.astype(np.float64) @ w.astype(np.float64) + b
    assert rel_err(hn, _ln_tanh(zv, sc.astype(np.float64), lb.astype(np.float64))[0]) < TOL
    mur, lsr = hn @ wm.astype(np.float64) + bm, hn @ wl.astype(np.float64) + bl
    assert rel_err(mu.cpu().numpy(), mur) < 1e-5 and rel_err(ls.cpu().numpy(), lsr) < 1e-5
    act2 = torch.zeros(M, A, device="cuda"); logp2 = torch.zeros(M, device="cuda"); u2 = torch.zeros(M, A, device="cuda"); sd2 = torch.zeros(M, A, device="cuda")
    ops.tanh_gaussian_fwd(mu, ls, t["eps"], 1e-5, 5.0, act2.data_ptr(), A, logp2, u2, sd2, M, A)
    np.testing.assert_allclose(act[:, :A].cpu().numpy(), act2.cpu().numpy(), rtol=0, atol=1e-6)
    np.testing.assert_allclose(logp.cpu().numpy(), logp2.cpu().numpy(), rtol=1e-5, atol=1e-5)
    np.testing.assert_array_equal(u.cpu().numpy(), u2.cpu().numpy())
    np.testing.assert_array_equal(sd.cpu().numpy(), sd2.cpu().numpy())
    assert float(act[:, A:].abs().max()) == 0.0                               # nothing written past column A


def test_rejects_unsupported_layouts():
    from serl_b200 import _lib as L
    from serl_b200 import ops
    x = torch.zeros(64, 7, device="cuda"); w = torch.zeros(7, 256, device="cuda"); c = torch.zeros(64, 256, device="cuda")
    p = ops.tgemm_problem(x.data_ptr(), w.data_ptr(), sAm=7, sAk=1, sBk=256, sBn=1, C_=c.data_ptr(), ldc=256)
    with pytest.raises(L.SerlError):
        ops.tgemm(None, [p], 64, 256, 7)                                     # row stride of 7 floats: not TMA-addressable
