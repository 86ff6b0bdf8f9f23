"""GPU: the batched CUDA-core companions of the TF32 head GEMMs (csrc/heads_fused.cu) one by one against float64 references -
several problems per launch, the shapes of the dual-camera critic step and ragged ones."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
TOL = 2e-5


def cu(x):
    return torch.as_tensor(np.ascontiguousarray(x)).cuda()


def rel_err(a, b):
    return float(np.max(np.abs(np.asarray(a, np.float64) - b)) / (np.max(np.abs(b)) + 1e-30))


def test_sle_fwd_multi_with_and_without_dropout():
    from serl_b200 import ops
    rng = np.random.default_rng(0)
    B = 37
    probs, refs, outs, keep = [], [], [], []
    for i in range(3):
        feat = np.abs(rng.standard_normal((B, 4, 4, 512))).astype(np.float32)
        kern = (rng.standard_normal((4, 4, 512, 8)) / 4).astype(np.float32)
        mask = (rng.random((B, 4096)) < 0.9).astype(np.uint8) if i == 2 else None
        ld = 4096 + (8 if i == 1 else 0)
        out = torch.full((B, ld), -7.0, device="cuda")
        t = (cu(feat), cu(kern), None if mask is None else cu(mask))
        keep.append(t)
        probs.append((t[0].data_ptr(), t[1].data_ptr(), None if mask is None else t[2].data_ptr(), out.data_ptr(), ld))
        ref = np.einsum("bhwc,hwcf->bcf", feat.astype(np.float64), kern.astype(np.float64)).reshape(B, 4096)
        if mask is not None:
            ref = np.where(mask.astype(bool), ref / 0.9, 0.0)
        refs.append(ref); outs.append((out, ld))
    ops.sle_fwd_multi(probs, 0.9, B, 16, 512)
    for (out, ld), ref in zip(outs, refs):
        o = out.cpu().numpy()
        assert rel_err(o[:, :4096], ref) < TOL
        assert (o[:, 4096:] == -7.0).all()


def _ln_tanh(z, sc, lb, eps=1e-6):
    mean = z.mean(-1, keepdims=True)
    var = np.maximum((z * z).mean(-1, keepdims=True) - mean * mean, 0.0)
    rstd = 1.0 / np.sqrt(var + eps)
    xh = (z - mean) * rstd
    return np.tanh(xh * sc + lb), xh, rstd[..., 0]


def test_enc_finish_partials_and_small_dense():
    from serl_b200 import ops
    rng = np.random.default_rng(1)
    rows, S = 70, 5
    part = rng.standard_normal((2, S, rows, 256)).astype(np.float32)
    b, sc, lb = [(0.2 * rng.standard_normal((2, 256))).astype(np.float32) for _ in range(3)]
    sc += 1
    x = rng.standard_normal((rows, 7)).astype(np.float32); w = rng.standard_normal((7, 64)).astype(np.float32)
    bp, scp, lbp = [(0.2 * rng.standard_normal(64)).astype(np.float32) for _ in range(3)]
    t = {k: cu(v) for k, v in dict(part=part, b=b, sc=sc, lb=lb, x=x, w=w, bp=bp, scp=scp, lbp=lbp).items()}
    X = torch.zeros(2, rows, 580, device="cuda"); xh = torch.zeros(rows, 256, device="cuda"); rs = torch.zeros(rows, device="cuda")
    xhp = torch.zeros(rows, 64, device="cuda"); rsp = torch.zeros(rows, device="cuda")
    probs = [dict(partials=t["part"].data_ptr() + 4 * i * S * rows * 256, S=S, bias=t["b"].data_ptr() + 1024 * i, ln_scale=t["sc"].data_ptr() + 1024 * i,
                  ln_bias=t["lb"].data_ptr() + 1024 * i, out=X.data_ptr() + 4 * i * rows * 580 + 4 * 256 * i, ld_out=580, D=256,
                  xhat=xh.data_ptr() if i == 0 else None, rstd=rs.data_ptr() if i == 0 else None) for i in range(2)]
    probs.append(dict(x=t["x"].data_ptr(), ld_x=7, w=t["w"].data_ptr(), K=7, bias=t["bp"].data_ptr(), ln_scale=t["scp"].data_ptr(), ln_bias=t["lbp"].data_ptr(),
                      out=X.data_ptr() + 4 * 512, ld_out=580, D=64, xhat=xhp.data_ptr(), rstd=rsp.data_ptr()))
    ops.enc_finish(probs, rows)
    Xh = X.cpu().numpy()
    for i in range(2):
        h, xr, rr = _ln_tanh(part[i].astype(np.float64).sum(0) + b[i], sc[i].astype(np.float64), lb[i].astype(np.float64))
        assert rel_err(Xh[i][:, 256 * i:256 * i + 256], h) < TOL
        if i == 0:
            assert rel_err(xh.cpu().numpy(), xr) < TOL and rel_err(rs.cpu().numpy(), rr) < TOL
    h, xr, rr = _ln_tanh(x.astype(np.float64) @ w.astype(np.float64) + bp, scp.astype(np.float64), lbp.astype(np.float64))
    assert rel_err(Xh[0][:, 512:576], h) < TOL and rel_err(xhp.cpu().numpy(), xr) < TOL and rel_err(rsp.cpu().numpy(), rr) < TOL
    assert (Xh[0][:, 576:] == 0).all()


def test_ln_tanh_bwd_multi_matches_single_kernel_and_outer_product_source():
    from serl_b200 import ops
    rng = np.random.default_rng(2)
    E, B, D = 3, 21, 256
    R = E * B
    t = np.tanh(rng.standard_normal((R, D))).astype(np.float32); xh = rng.standard_normal((R, D)).astype(np.float32)
    rstd = (0.5 + rng.random(R)).astype(np.float32); sc = (1 + 0.1 * rng.standard_normal((E, D))).astype(np.float32)
    dq = rng.standard_normal(R).astype(np.float32); hw = rng.standard_normal(D).astype(np.float32)
    parts = rng.standard_normal((4, R, D)).astype(np.float32)
    T = {k: cu(v) for k, v in dict(t=t, xh=xh, rstd=rstd, sc=sc, dq=dq, hw=hw, parts=parts).items()}

    def ref(dt):
        dy = dt * (1 - t.astype(np.float64) ** 2)
        dxh = dy * np.repeat(sc.astype(np.float64), B, axis=0)
        m1, m2 = dxh.mean(-1, keepdims=True), (dxh * xh).mean(-1, keepdims=True)
        return rstd[:, None] * (dxh - m1 - xh * m2), dy

    dz1, dy1, dz2 = (torch.zeros(R, D, device="cuda") for _ in range(3))
    base = dict(t=T["t"].data_ptr(), ld_t=D, xhat=T["xh"].data_ptr(), rstd=T["rstd"].data_ptr(), scale=T["sc"].data_ptr(), rows_per_group=B, group_stride=D, R=R, D=D)
    ops.ln_tanh_bwd_multi([dict(base, dq=T["dq"].data_ptr(), head_w=T["hw"].data_ptr(), dz=dz1.data_ptr(), dy=dy1.data_ptr()),
                           dict(base, dt=T["parts"].data_ptr(), ld_dt=D, dt_parts=4, dt_part_stride=R * D, dz=dz2.data_ptr())])
    rz, ry = ref(np.outer(dq.astype(np.float64), hw.astype(np.float64)))
    assert rel_err(dz1.cpu().numpy(), rz) < TOL and rel_err(dy1.cpu().numpy(), ry) < TOL
    rz2, _ = ref(parts.astype(np.float64).sum(0))
    assert rel_err(dz2.cpu().numpy(), rz2) < TOL


def test_small_grads_and_sle_bwd_multi():
    from serl_b200 import _lib as L
    from serl_b200 import ops
    rng = np.random.default_rng(3)
    E, B, D = 4, 50, 256
    x = rng.standard_normal((E * B, D)).astype(np.float32); y = rng.standard_normal((E * B, D)).astype(np.float32)
    dq = rng.standard_normal(E * B).astype(np.float32); xs = rng.standard_normal((B, 64)).astype(np.float32)
    T = {k: cu(v) for k, v in dict(x=x, y=y, dq=dq, xs=xs).items()}
    cs = torch.zeros(E, D, device="cuda"); la = torch.zeros(E, D, device="cuda"); lb = torch.zeros(E, D, device="cuda")
    hw = torch.zeros(D, device="cuda"); hb = torch.zeros(1, device="cuda"); c64 = torch.zeros(64, device="cuda")
    ops.small_grads([(L.SMALL_GRAD_COLSUM, T["x"].data_ptr(), D, None, 0, cs.data_ptr(), None, E, B, D),
                     (L.SMALL_GRAD_LN, T["x"].data_ptr(), D, T["y"].data_ptr(), D, la.data_ptr(), lb.data_ptr(), E, B, D),
                     (L.SMALL_GRAD_HEAD, T["x"].data_ptr(), D, T["dq"].data_ptr(), 1, hw.data_ptr(), hb.data_ptr(), 1, E * B, D),
                     (L.SMALL_GRAD_COLSUM, T["xs"].data_ptr(), 64, None, 0, c64.data_ptr(), None, 1, B, 64)])
    x64, y64 = x.astype(np.float64).reshape(E, B, D), y.astype(np.float64).reshape(E, B, D)
    assert rel_err(cs.cpu().numpy(), x64.sum(1)) < TOL and rel_err(la.cpu().numpy(), (x64 * y64).sum(1)) < TOL and rel_err(lb.cpu().numpy(), x64.sum(1)) < TOL
    assert rel_err(hw.cpu().numpy(), (x.astype(np.float64) * dq[:, None]).sum(0)) < TOL and abs(float(hb.item()) - dq.astype(np.float64).sum()) < 1e-4
    assert rel_err(c64.cpu().numpy(), xs.astype(np.float64).sum(0)) < TOL
    # SLE kernel gradients of two cameras in one call == the single-problem entry point
    N = 70
    ws = ops.Workspace(64 << 20, "cuda")
    feats = [cu(np.abs(rng.standard_normal((N, 4, 4, 512))).astype(np.float32)) for _ in range(2)]
    douts = [cu(rng.standard_normal((N, 4096)).astype(np.float32)) for _ in range(2)]
    outs = [torch.zeros(4, 4, 512, 8, device="cuda") for _ in range(2)]
    ops.sle_bwd_multi(ws, [(f.data_ptr(), d.data_ptr(), 4096, o.data_ptr()) for f, d, o in zip(feats, douts, outs)], N, 16, 512)
    for f, d, o in zip(feats, douts, outs):
        ref = np.einsum("bhwc,bcf->hwcf", f.cpu().numpy().astype(np.float64), d.cpu().numpy().astype(np.float64).reshape(N, 512, 8))
        assert rel_err(o.cpu().numpy(), ref) < TOL
