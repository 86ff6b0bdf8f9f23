"""Device-side execution plan of one DrQ/SAC `update` call: which hand-written kernels run, in what
order, on which HBM buffers.  Pure orchestration - every arithmetic op is a C-ABI call (ops.py).

Data flow (pixel agent, one camera shown; B = batch, N = 2B images):

  sampler kernel ─ u8 crops (obs rows [0,B), next rows [B,2B)) ─ trunk (frozen ResNet-10, ONCE per step,
  shared by policy / critic / target critic; the reference recomputes it per network)
  ─ feats (N,4,4,512) ─ trainable heads (SLE, Dropout, Dense, LN, tanh; + proprio) ─ enc (B,F)
  ─ policy MLP / critic ensemble ─ losses ─ analytic backward ─ flat gradient ─ fused Adam+polyak.

Reference semantics: agents/continuous/sac.py:134-299, agents/continuous/drq.py:255-328,
common/common.py:124-221 (see oracle/drq.py for the restatement this is tested against).
"""
from __future__ import annotations

import os

from dataclasses import dataclass
from typing import Dict, Optional, Sequence

import torch

from . import _lib as L
from . import ops
from .params import ENC, INFO_GAP, STAGES, ParamStore

f32 = torch.float32


@dataclass
class AgentConfig:
    cams: Sequence[str]
    state_in: int                # T * S (pixel agent) or S (state agent)
    action_dim: int
    pixel: bool = True
    ensemble: int = 10
    subsample: Optional[int] = 2
    discount: float = 0.96
    tau: float = 0.005
    target_entropy: float = -2.0
    backup_entropy: bool = False
    lr: Sequence[float] = (3e-4, 3e-4, 3e-4)            # critic, actor, temperature tx
    warmup: Sequence[int] = (0, 0, 0)
    std_min: float = 1e-5
    std_max: float = 5.0
    image_hw: int = 128
    precision: str = "fp32"      # trunk arithmetic: "fp32" (1e-5 parity build) | "bf16" / "fp16" (tcgen05 builds)

    @property
    def enc_dim(self):
        return 256 * len(self.cams) + 64 if self.pixel else self.state_in


class _MlpActs:
    """Activations of a 2-layer Dense-LN-tanh stack over R rows (R = E*B for the ensemble)."""

    def __init__(self, R, dev, H=256):
        e = lambda *s: torch.empty(*s, dtype=f32, device=dev)
        self.z = e(R, H)                           # pre-LN scratch (reused by both layers)
        self.h1, self.xhat1, self.rstd1 = e(R, H), e(R, H), e(R)
        self.h2, self.xhat2, self.rstd2 = e(R, H), e(R, H), e(R)


class _EncScratch:
    """Scratch of one encoder-heads pass; each concurrently running branch of the step owns one."""

    def __init__(self, cfg, B, device, ws):
        e = lambda *s: torch.empty(*s, dtype=f32, device=device)
        self.ws = ws
        if cfg.pixel:
            self.sle = {c: e(B, 4096) for c in cfg.cams}
            self.enc_z, self.enc_zp = e(B, 256), e(B, 64)


class Engine:
    def __init__(self, cfg: AgentConfig, store: ParamStore, trunk: Dict[str, Dict[str, torch.Tensor]], batch: int, device):
        self.cfg, self.store, self.trunk, self.B, self.dev = cfg, store, trunk, batch, device
        B, E, A, F = batch, cfg.ensemble, cfg.action_dim, cfg.enc_dim
        self.F, self.FA = F, F + A
        e = lambda *s: torch.empty(*s, dtype=f32, device=device)
        # heads GEMMs: CUDA-core SGEMM in the 1e-5 build, tensor-core 3xTF32 (fp32-class accuracy) next to the 16-bit trunk
        gemm_impl = os.environ.get("SERL_HEADS_GEMM") or ("f32" if cfg.precision == "fp32" else "tf32x3")
        ws_bytes = max(48 << 20, 2 * 4 * E * B * self.FA)
        self.ws = ops.Workspace(ws_bytes, device, gemm_impl)
        # Branch-level concurrency: the heads are ~100 short, latency-bound launches, and several chains of them are
        # independent (online critic / target encoder / policy in the forward pass; weight gradients vs the dX chain in
        # the backward pass).  They run on two side streams; under CUDA-graph capture the fork / join events become graph
        # edges.  Every branch owns its scratch (split-K workspace included).
        dev = torch.device(device)
        streams_on = os.environ.get("SERL_STREAMS", "1") != "0"
        self.side = [L.new_side_stream(dev, streams_on) for _ in range(2)]
        # frozen trunk: every camera's pass on its own stream (camera 0 stays on the main stream), each with its own side stream
        # for the block's projection conv.  At batch 256 the persistent conv kernels fill the GPU and the passes serialise; at
        # the small per-rank batches of data-parallel runs (32 rows per rank on 8 GPUs) a trunk kernel covers a fraction of the
        # SMs and the cameras overlap.
        self.cam_stream = {c: (L.new_side_stream(dev, streams_on and os.environ.get("SERL_CAM_STREAMS", "1") != "0") if j > 0 else None)
                           for j, c in enumerate(cfg.cams)}
        self.proj_side = {c: L.new_side_stream(dev, streams_on and os.environ.get("SERL_PROJ_SIDE", "1") != "0") for c in cfg.cams}
        self.ws_side = [ops.Workspace(ws_bytes, device, gemm_impl) for _ in range(2)]
        # batch tensors
        self.state_o, self.state_n = e(B, cfg.state_in), e(B, cfg.state_in)
        self.actions, self.rewards, self.masks = e(B, A), e(B), e(B)
        self.dones = torch.empty(B, dtype=torch.uint8, device=device)
        self.idx = torch.empty(B, dtype=torch.int32, device=device)
        self.status = torch.zeros(1, dtype=torch.int32, device=device)
        if cfg.pixel:
            hw, N = cfg.image_hw, 2 * B
            self.N = N
            self.pix = {c: torch.empty(N, hw, hw, 3, dtype=torch.uint8, device=device) for c in cfg.cams}
            self.off = torch.empty(2, B, 2, dtype=torch.int32, device=device)       # applied crop offsets (obs, next)
            self.feats = {c: e(N, 4, 4, 512) for c in cfg.cams}
            s2 = hw // 2
            self.t_a0 = e(N, s2, s2, 64)
            self.t_buf = [e(N * (s2 // 2) * (s2 // 2) * 64) for _ in range(4)]
            self.sle_saved = {c: e(B, 4096) for c in cfg.cams}
            self.enc_xhat = {c: e(B, 256) for c in cfg.cams}
            self.enc_rstd = {c: e(B) for c in cfg.cams}
            self.enc_xhat_p, self.enc_rstd_p = e(B, 64), e(B)
            self.masks_u8 = {c: torch.empty(B, 4096, dtype=torch.uint8, device=device) for c in cfg.cams}
            self.d_enc_z = {c: e(B, 256) for c in cfg.cams}          # per camera: the side stream reads it while the next one is written
            self.d_enc_y, self.d_sle = {c: e(B, 256) for c in cfg.cams}, e(B, 4096)
            self.d_enc_zp, self.d_enc_yp = e(B, 64), e(B, 64)
            # actor pass: the proprio Dense/LayerNorm stays differentiable under the policy's stop_gradient (encoding.py:48-70)
            self.enc_xhat_pa, self.enc_rstd_pa = e(B, 64), e(B)
            self.dXp_p, self.d_enc_zpa, self.d_enc_ypa = e(B, 64), e(B, 64), e(B, 64)
        self.sc_main = _EncScratch(cfg, B, device, self.ws)
        self.sc_side = [_EncScratch(cfg, B, device, w) for w in self.ws_side]
        # critic / policy activations
        self.Xc, self.Xt, self.Xp = e(B, self.FA), e(B, self.FA), e(B, F)
        self.c_main, self.c_tgt = _MlpActs(E * B, device), _MlpActs(E * B, device)
        self.q, self.q_next, self.dq, self.target_q = e(E, B), e(E, B), e(E, B), e(B)
        self.p_acts = _MlpActs(B, device)
        self.mu, self.ls, self.u, self.std, self.eps = e(B, A), e(B, A), e(B, A), e(B, A), e(B, A)
        self.logp = e(B)
        self.act_scratch = e(B, A)
        self.sub = torch.zeros(max(cfg.subsample or 1, 1), dtype=torch.int32, device=device)
        # gradient scratch
        self.dh, self.dz, self.dy = e(E * B, 256), e(E * B, 256), e(E * B, 256)
        self.dz0, self.dy0 = e(E * B, 256), e(E * B, 256)           # layer-0 dz / dy: layer-1's are still read by the side stream
        self.dX = e(B, self.FA)
        self.dmu, self.dls = e(B, A), e(B, A)
        self.pdh, self.pdz, self.pdy = e(B, 256), e(B, 256), e(B, 256)
        # info scalars live INSIDE the flat gradient buffer (params.py: info gap), next to the segments they travel with in
        # the data-parallel all-reduce: [0:3] critic | [4:7] actor, [8] temperature.  Learning rates are separate.
        self.info = store.grad[store.info_off:store.info_off + INFO_GAP]
        self.info.zero_()
        self.info_hist = torch.zeros(INFO_GAP, dtype=f32, device=device)
        self.lr_info = torch.zeros(4, dtype=f32, device=device)
        self.launches = 0
        # 16-bit builds, pixel agent: the critic step runs on the fused head kernels (heads_fused.py: TF32 GEMMs with TMA-fed
        # operands and LayerNorm / head epilogues, batched problems); SERL_FUSED_HEADS=0 keeps the per-op chain below.
        from . import heads_fused
        self.fused = heads_fused.FusedCritic(self) if (heads_fused.enabled(cfg) and (dev.type == "cuda" or os.environ.get("SERL_FUSED_HEADS") == "force")) else None

    # ------------------------------------------------------------------------------------------
    def P(self, buf, path):
        return self.store.addr(buf, path)

    # ---- frozen trunk (vision/resnet_v1.py:217-286) -------------------------------------------
    def trunk_forward(self, cam: str, pix: torch.Tensor, feats: torch.Tensor):
        if self.cfg.precision != "fp32":
            from . import trunk_bf16
            return trunk_bf16.forward(self, cam, pix, feats)
        w = self.trunk[cam]
        N, hw = pix.shape[0], pix.shape[1]
        s = hw // 2
        a0 = self.t_a0[:N]
        ops.conv2d_nhwc(pix, w["conv_init/kernel"], a0, 2, 3, 3)
        ops.groupnorm_nhwc(a0, a0, w["norm_init/scale"], w["norm_init/bias"], None, 4, 1e-5, True)
        s //= 2
        x = self.t_buf[0][:N * s * s * 64].view(N, s, s, 64)
        ops.maxpool3x3s2_nhwc(a0, x)
        free = [1, 2, 3]
        cur = 0
        cin = 64
        for i, (f, stride) in enumerate(STAGES):
            b = f"ResNetBlock_{i}"
            so = s // stride
            iy, iy2, ir = free
            y = self.t_buf[iy][:N * so * so * f].view(N, so, so, f)
            lo, hi = (1, 1) if stride == 1 else (0, 1)           # XLA SAME on even sizes
            ops.conv2d_nhwc(x, w[f"{b}/Conv_0/kernel"], y, stride, lo, hi)
            ops.groupnorm_nhwc(y, y, w[f"{b}/MyGroupNorm_0/scale"], w[f"{b}/MyGroupNorm_0/bias"], None, 4, 1e-5, True)
            last = i == len(STAGES) - 1
            y2 = feats[:N] if last else self.t_buf[iy2][:N * so * so * f].view(N, so, so, f)
            ops.conv2d_nhwc(y, w[f"{b}/Conv_1/kernel"], y2, 1, 1, 1)
            if stride != 1 or cin != f:
                r = self.t_buf[ir][:N * so * so * f].view(N, so, so, f)
                ops.conv2d_nhwc(x, w[f"{b}/conv_proj/kernel"], r, stride, 0, 0)
                ops.groupnorm_nhwc(r, r, w[f"{b}/norm_proj/scale"], w[f"{b}/norm_proj/bias"], None, 4, 1e-5, False)
                self.launches += 2
            else:
                r = x
            ops.groupnorm_nhwc(y2, y2, w[f"{b}/MyGroupNorm_1/scale"], w[f"{b}/MyGroupNorm_1/bias"], r, 4, 1e-5, True)
            self.launches += 4
            if not last:
                free = [cur, iy, ir]
                cur = iy2
                x, s, cin = y2, so, f
        self.launches += 3
        return feats

    # ---- trainable encoder heads (common/encoding.py:26-72, vision/resnet_v1.py:340-374) -------
    def encode(self, buf, feats_rows: slice, state: torch.Tensor, out: torch.Tensor, ld_out: int,
               masks: Optional[Dict[str, torch.Tensor]], save: bool, sc: Optional[_EncScratch] = None, save_proprio_actor: bool = False):
        """save: keep what the critic-loss backward needs (every head).  save_proprio_actor: keep the proprio LayerNorm
        statistics for the ACTOR-loss backward (the only encoder branch the policy's stop_gradient leaves differentiable)."""
        sc = sc or self.sc_main
        cfg, B, ws = self.cfg, self.B, sc.ws
        if not cfg.pixel:
            ops.copy2d(state.data_ptr(), cfg.state_in, out.data_ptr(), ld_out, B, cfg.state_in)
            self.launches += 1
            return
        for j, cam in enumerate(cfg.cams):
            p = f"{ENC}/encoder_{cam}"
            sle = self.sle_saved[cam] if save else sc.sle[cam]
            ops.sle_fwd(self.feats[cam][feats_rows], self.store.view(buf, f"{p}/SpatialLearnedEmbeddings_0/kernel"),
                        None if masks is None else masks[cam], 0.9, sle.data_ptr(), 4096)
            ops.dense_fwd(ws, sle.data_ptr(), 4096, self.P(buf, f"{p}/Dense_0/kernel"), self.P(buf, f"{p}/Dense_0/bias"),
                          sc.enc_z.data_ptr(), 256, B, 4096, 256)
            ops.ln_tanh_fwd(sc.enc_z.data_ptr(), 256, self.P(buf, f"{p}/LayerNorm_0/scale"), self.P(buf, f"{p}/LayerNorm_0/bias"),
                            B, 0, ops.at(out, 256 * j), ld_out, self.enc_xhat[cam].data_ptr() if save else None,
                            self.enc_rstd[cam].data_ptr() if save else None, B, 256)
            self.launches += 4
        ops.dense_fwd(ws, state.data_ptr(), cfg.state_in, self.P(buf, f"{ENC}/Dense_0/kernel"), self.P(buf, f"{ENC}/Dense_0/bias"),
                      sc.enc_zp.data_ptr(), 64, B, cfg.state_in, 64)
        xh, rs = (self.enc_xhat_p, self.enc_rstd_p) if save else ((self.enc_xhat_pa, self.enc_rstd_pa) if save_proprio_actor else (None, None))
        ops.ln_tanh_fwd(sc.enc_zp.data_ptr(), 64, self.P(buf, f"{ENC}/LayerNorm_0/scale"), self.P(buf, f"{ENC}/LayerNorm_0/bias"),
                        B, 0, ops.at(out, 256 * len(cfg.cams)), ld_out, None if xh is None else xh.data_ptr(),
                        None if rs is None else rs.data_ptr(), B, 64)
        self.launches += 2

    def encode_backward(self, dX: torch.Tensor, X: torch.Tensor, feats_rows: slice, state: torch.Tensor):
        """Gradients of the trainable heads given d(enc) = dX[:, :F]; trunk is stop-gradient.
        Weight / bias gradients run on side stream 0, the d_enc_z -> d_sle -> SLE-kernel chain stays on the main stream."""
        cfg, B, ws, st = self.cfg, self.B, self.ws, self.store
        G = st.grad
        ld = self.FA
        side, wss = self.side[0], self.ws_side[0]
        for j, cam in enumerate(cfg.cams):
            p = f"{ENC}/encoder_{cam}"
            dez = self.d_enc_z[cam]
            dey = self.d_enc_y[cam]
            ops.ln_tanh_bwd(ops.at(dX, 256 * j), ld, ops.at(X, 256 * j), ld, self.enc_xhat[cam].data_ptr(), self.enc_rstd[cam].data_ptr(),
                            self.P(st.params, f"{p}/LayerNorm_0/scale"), B, 0, dez.data_ptr(), dey.data_ptr(), None, None, B, 256)
            side.fork()
            with side:
                ops.ln_param_grad(dey.data_ptr(), self.enc_xhat[cam].data_ptr(), self.P(G, f"{p}/LayerNorm_0/scale"),
                                  self.P(G, f"{p}/LayerNorm_0/bias"), B, B, 256)
                ops.dense_bwd_weight(wss, self.sle_saved[cam].data_ptr(), 4096, dez.data_ptr(), 256, self.P(G, f"{p}/Dense_0/kernel"),
                                     B, 4096, 256)
                ops.colsum(dez.data_ptr(), self.P(G, f"{p}/Dense_0/bias"), 1, B, 256, 256)
            ops.dense_bwd_input(ws, dez.data_ptr(), 256, self.P(st.params, f"{p}/Dense_0/kernel"), self.d_sle.data_ptr(), 4096,
                                B, 4096, 256)
            ops.sle_bwd_kernel_grad(ws, self.feats[cam][feats_rows], self.d_sle.data_ptr(), 4096,
                                    self.P(G, f"{p}/SpatialLearnedEmbeddings_0/kernel"))
            self.launches += 9
        off = 256 * len(cfg.cams)
        ops.ln_tanh_bwd(ops.at(dX, off), ld, ops.at(X, off), ld, self.enc_xhat_p.data_ptr(), self.enc_rstd_p.data_ptr(),
                        self.P(st.params, f"{ENC}/LayerNorm_0/scale"), B, 0, self.d_enc_zp.data_ptr(), self.d_enc_yp.data_ptr(),
                        None, None, B, 64)
        side.fork()
        with side:
            ops.ln_param_grad(self.d_enc_yp.data_ptr(), self.enc_xhat_p.data_ptr(), self.P(G, f"{ENC}/LayerNorm_0/scale"),
                              self.P(G, f"{ENC}/LayerNorm_0/bias"), B, B, 64)
            ops.dense_bwd_weight(wss, state.data_ptr(), cfg.state_in, self.d_enc_zp.data_ptr(), 64, self.P(G, f"{ENC}/Dense_0/kernel"),
                                 B, cfg.state_in, 64)
            ops.colsum(self.d_enc_zp.data_ptr(), self.P(G, f"{ENC}/Dense_0/bias"), 1, B, 64, 64)
        self.launches += 4

    # ---- critic ensemble (networks/actor_critic_nets.py:57-73, networks/mlp.py:22-31) ----------
    def critic_forward(self, buf, X: torch.Tensor, acts: _MlpActs, q: torch.Tensor, save: bool, ws: Optional[ops.Workspace] = None):
        cfg, B, E, ws = self.cfg, self.B, self.cfg.ensemble, ws or self.ws
        c = "modules_critic/network"
        FA = self.FA
        ops.dense_fwd(ws, X.data_ptr(), FA, self.P(buf, f"{c}/Dense_0/kernel"), self.P(buf, f"{c}/Dense_0/bias"), acts.z.data_ptr(), 256,
                      B, FA, 256, Z=E, x_z=0, out_z=B * 256)
        ops.ln_tanh_fwd(acts.z.data_ptr(), 256, self.P(buf, f"{c}/LayerNorm_0/scale"), self.P(buf, f"{c}/LayerNorm_0/bias"), B, 256,
                        acts.h1.data_ptr(), 256, acts.xhat1.data_ptr() if save else None, acts.rstd1.data_ptr() if save else None, E * B, 256)
        ops.dense_fwd(ws, acts.h1.data_ptr(), 256, self.P(buf, f"{c}/Dense_1/kernel"), self.P(buf, f"{c}/Dense_1/bias"), acts.z.data_ptr(), 256,
                      B, 256, 256, Z=E, x_z=B * 256, out_z=B * 256)
        ops.ln_tanh_fwd(acts.z.data_ptr(), 256, self.P(buf, f"{c}/LayerNorm_1/scale"), self.P(buf, f"{c}/LayerNorm_1/bias"), B, 256,
                        acts.h2.data_ptr(), 256, acts.xhat2.data_ptr() if save else None, acts.rstd2.data_ptr() if save else None, E * B, 256)
        wk, wb = self.P(buf, "modules_critic/Dense_0/kernel"), self.P(buf, "modules_critic/Dense_0/bias")
        if cfg.pixel:     # one shared head over all E*B rows
            ops.dense_fwd(ws, acts.h2.data_ptr(), 256, wk, wb, q.data_ptr(), 1, E * B, 256, 1)
        else:             # per-member head
            ops.dense_fwd(ws, acts.h2.data_ptr(), 256, wk, wb, q.data_ptr(), 1, B, 256, 1, Z=E, x_z=B * 256, w_z=256, b_z=1, out_z=B)
        self.launches += 7

    def critic_backward(self, X: torch.Tensor, acts: _MlpActs, dq: torch.Tensor, param_grads: bool, need_dx: bool):
        """The dq -> dh -> dz -> ... -> dX chain runs on the main stream; each layer's weight / bias gradient only needs that
        layer's (input, dz) pair, so it is forked to side stream 0 as soon as dz exists (joined by the caller)."""
        cfg, B, E, ws, st = self.cfg, self.B, self.cfg.ensemble, self.ws, self.store
        G, Pm = st.grad, st.params
        c = "modules_critic/network"
        FA, R = self.FA, E * B
        dh, dz, dz0, dy, dy0 = self.dh.data_ptr(), self.dz.data_ptr(), self.dz0.data_ptr(), self.dy.data_ptr(), self.dy0.data_ptr()
        side, wss = self.side[0], self.ws_side[0]
        wk = self.P(Pm, "modules_critic/Dense_0/kernel")
        if param_grads:
            side.fork()
            with side:
                if cfg.pixel:
                    ops.dense_bwd_weight(wss, acts.h2.data_ptr(), 256, dq.data_ptr(), 1, self.P(G, "modules_critic/Dense_0/kernel"), R, 256, 1)
                    ops.colsum(dq.data_ptr(), self.P(G, "modules_critic/Dense_0/bias"), 1, R, 1, 1)
                else:
                    ops.dense_bwd_weight(wss, acts.h2.data_ptr(), 256, dq.data_ptr(), 1, self.P(G, "modules_critic/Dense_0/kernel"), B, 256, 1,
                                         Z=E, x_z=B * 256, dz_z=B, dw_z=256)
                    ops.colsum(dq.data_ptr(), self.P(G, "modules_critic/Dense_0/bias"), E, B, 1, 1)
        if cfg.pixel:
            ops.dense_bwd_input(ws, dq.data_ptr(), 1, wk, dh, 256, R, 256, 1)
        else:
            ops.dense_bwd_input(ws, dq.data_ptr(), 1, wk, dh, 256, B, 256, 1, Z=E, dz_z=B, w_z=256, dx_z=B * 256)
        ops.ln_tanh_bwd(dh, 256, acts.h2.data_ptr(), 256, acts.xhat2.data_ptr(), acts.rstd2.data_ptr(), self.P(Pm, f"{c}/LayerNorm_1/scale"), B, 256,
                        dz, dy, None, None, R, 256)
        if param_grads:
            side.fork()
            with side:
                ops.ln_param_grad(dy, acts.xhat2.data_ptr(), self.P(G, f"{c}/LayerNorm_1/scale"), self.P(G, f"{c}/LayerNorm_1/bias"), B, R, 256)
                ops.dense_bwd_weight(wss, acts.h1.data_ptr(), 256, dz, 256, self.P(G, f"{c}/Dense_1/kernel"), B, 256, 256, Z=E, x_z=B * 256, dz_z=B * 256)
                ops.colsum(dz, self.P(G, f"{c}/Dense_1/bias"), E, B, 256, 256)
        ops.dense_bwd_input(ws, dz, 256, self.P(Pm, f"{c}/Dense_1/kernel"), dh, 256, B, 256, 256, Z=E, dz_z=B * 256, dx_z=B * 256)
        ops.ln_tanh_bwd(dh, 256, acts.h1.data_ptr(), 256, acts.xhat1.data_ptr(), acts.rstd1.data_ptr(), self.P(Pm, f"{c}/LayerNorm_0/scale"), B, 256,
                        dz0, dy0, None, None, R, 256)
        if param_grads:
            side.fork()
            with side:
                ops.ln_param_grad(dy0, acts.xhat1.data_ptr(), self.P(G, f"{c}/LayerNorm_0/scale"), self.P(G, f"{c}/LayerNorm_0/bias"), B, R, 256)
                ops.dense_bwd_weight(wss, X.data_ptr(), FA, dz0, 256, self.P(G, f"{c}/Dense_0/kernel"), B, FA, 256, Z=E, x_z=0, dz_z=B * 256)
                ops.colsum(dz0, self.P(G, f"{c}/Dense_0/bias"), E, B, 256, 256)
        if need_dx:       # input is broadcast over the ensemble: dX = sum_e dZ1_e W1_e^T
            ops.dense_bwd_input(ws, dz0, 256, self.P(Pm, f"{c}/Dense_0/kernel"), self.dX.data_ptr(), FA, B, FA, 256, Z=E, dz_z=B * 256,
                                reduce_z=True)
        self.launches += 8 + (7 if param_grads else 0) + (2 if need_dx else 0)

    # ---- policy (networks/actor_critic_nets.py:178-227) ------------------------------------------
    def policy_forward(self, buf, Xp: torch.Tensor, save: bool):
        B, ws, a, A = self.B, self.ws, self.p_acts, self.cfg.action_dim
        n = "modules_actor/network"
        F = self.F
        ops.dense_fwd(ws, Xp.data_ptr(), F, self.P(buf, f"{n}/Dense_0/kernel"), self.P(buf, f"{n}/Dense_0/bias"), a.z.data_ptr(), 256, B, F, 256)
        ops.ln_tanh_fwd(a.z.data_ptr(), 256, self.P(buf, f"{n}/LayerNorm_0/scale"), self.P(buf, f"{n}/LayerNorm_0/bias"), B, 0,
                        a.h1.data_ptr(), 256, a.xhat1.data_ptr() if save else None, a.rstd1.data_ptr() if save else None, B, 256)
        ops.dense_fwd(ws, a.h1.data_ptr(), 256, self.P(buf, f"{n}/Dense_1/kernel"), self.P(buf, f"{n}/Dense_1/bias"), a.z.data_ptr(), 256, B, 256, 256)
        ops.ln_tanh_fwd(a.z.data_ptr(), 256, self.P(buf, f"{n}/LayerNorm_1/scale"), self.P(buf, f"{n}/LayerNorm_1/bias"), B, 0,
                        a.h2.data_ptr(), 256, a.xhat2.data_ptr() if save else None, a.rstd2.data_ptr() if save else None, B, 256)
        ops.dense_fwd(ws, a.h2.data_ptr(), 256, self.P(buf, "modules_actor/Dense_0/kernel"), self.P(buf, "modules_actor/Dense_0/bias"),
                      self.mu.data_ptr(), A, B, 256, A)
        ops.dense_fwd(ws, a.h2.data_ptr(), 256, self.P(buf, "modules_actor/Dense_1/kernel"), self.P(buf, "modules_actor/Dense_1/bias"),
                      self.ls.data_ptr(), A, B, 256, A)
        self.launches += 8

    def policy_backward(self, Xp: torch.Tensor):
        B, ws, a, A, st = self.B, self.ws, self.p_acts, self.cfg.action_dim, self.store
        G, Pm = st.grad, st.params
        n = "modules_actor/network"
        F = self.F
        dmu, dls = self.dmu.data_ptr(), self.dls.data_ptr()
        dh, dz, dy = self.pdh.data_ptr(), self.pdz.data_ptr(), self.pdy.data_ptr()
        ops.dense_bwd_weight(ws, a.h2.data_ptr(), 256, dmu, A, self.P(G, "modules_actor/Dense_0/kernel"), B, 256, A)
        ops.colsum(dmu, self.P(G, "modules_actor/Dense_0/bias"), 1, B, A, A)
        ops.dense_bwd_weight(ws, a.h2.data_ptr(), 256, dls, A, self.P(G, "modules_actor/Dense_1/kernel"), B, 256, A)
        ops.colsum(dls, self.P(G, "modules_actor/Dense_1/bias"), 1, B, A, A)
        ops.dense_bwd_input(ws, dmu, A, self.P(Pm, "modules_actor/Dense_0/kernel"), dh, 256, B, 256, A)
        ops.dense_bwd_input(ws, dls, A, self.P(Pm, "modules_actor/Dense_1/kernel"), dh, 256, B, 256, A, accumulate=True)
        ops.ln_tanh_bwd(dh, 256, a.h2.data_ptr(), 256, a.xhat2.data_ptr(), a.rstd2.data_ptr(), self.P(Pm, f"{n}/LayerNorm_1/scale"), B, 0,
                        dz, dy, self.P(G, f"{n}/LayerNorm_1/scale"), self.P(G, f"{n}/LayerNorm_1/bias"), B, 256)
        ops.dense_bwd_weight(ws, a.h1.data_ptr(), 256, dz, 256, self.P(G, f"{n}/Dense_1/kernel"), B, 256, 256)
        ops.colsum(dz, self.P(G, f"{n}/Dense_1/bias"), 1, B, 256, 256)
        ops.dense_bwd_input(ws, dz, 256, self.P(Pm, f"{n}/Dense_1/kernel"), dh, 256, B, 256, 256)
        ops.ln_tanh_bwd(dh, 256, a.h1.data_ptr(), 256, a.xhat1.data_ptr(), a.rstd1.data_ptr(), self.P(Pm, f"{n}/LayerNorm_0/scale"), B, 0,
                        dz, dy, self.P(G, f"{n}/LayerNorm_0/scale"), self.P(G, f"{n}/LayerNorm_0/bias"), B, 256)
        ops.dense_bwd_weight(ws, Xp.data_ptr(), F, dz, 256, self.P(G, f"{n}/Dense_0/kernel"), B, F, 256)
        ops.colsum(dz, self.P(G, f"{n}/Dense_0/bias"), 1, B, 256, 256)
        self.launches += 17
        if self.cfg.pixel:
            # Policy.__call__ -> encoder(..., stop_gradient=True) (actor_critic_nets.py:185) stops the gradient at the per-camera
            # image embeddings only (encoding.py:48-49); the proprio Dense -> LayerNorm -> tanh (:55-70) is differentiated by
            # jax.grad(policy_loss_fn) w.r.t. the full tree (sac.py:198-200).  Its gradient goes to the ACTOR-tx twin (aux tail)
            # of those leaves: d enc[:, off:] = dz0 @ W0[off:, :]^T, then back through LayerNorm / tanh / Dense.
            off, S = 256 * len(self.cfg.cams), self.cfg.state_in
            ops.dense_bwd_input(ws, dz, 256, self.P(Pm, f"{n}/Dense_0/kernel") + 4 * off * 256, self.dXp_p.data_ptr(), 64, B, 64, 256)
            ops.ln_tanh_bwd(self.dXp_p.data_ptr(), 64, ops.at(Xp, off), F, self.enc_xhat_pa.data_ptr(), self.enc_rstd_pa.data_ptr(),
                            self.P(Pm, f"{ENC}/LayerNorm_0/scale"), B, 0, self.d_enc_zpa.data_ptr(), self.d_enc_ypa.data_ptr(),
                            st.aux_addr(G, f"{ENC}/LayerNorm_0/scale"), st.aux_addr(G, f"{ENC}/LayerNorm_0/bias"), B, 64)
            ops.dense_bwd_weight(ws, self.pol_state.data_ptr(), S, self.d_enc_zpa.data_ptr(), 64, st.aux_addr(G, f"{ENC}/Dense_0/kernel"), B, S, 64)
            ops.colsum(self.d_enc_zpa.data_ptr(), st.aux_addr(G, f"{ENC}/Dense_0/bias"), 1, B, 64, 64)
            self.launches += 4

    # ---- the three losses --------------------------------------------------------------------------
    def _policy_pass(self, feats_rows, state, key_slot_eps, key_slot_drop, keys, act_out, ld_act, save, explicit=None):
        """enc(train=True -> dropout) -> policy -> tanh-Gaussian sample.  Writes actions to act_out."""
        cfg, B, A, st = self.cfg, self.B, self.cfg.action_dim, self.store
        if explicit is None:
            ops.normal_fill(ops.key_ptr(keys, key_slot_eps), self.eps, B * A)
            self.launches += 1
            if cfg.pixel:
                for j, cam in enumerate(cfg.cams):
                    ops.dropout_mask_fill(ops.key_ptr(keys, key_slot_drop), j, 0.9, self.masks_u8[cam], B * 4096)
                    self.launches += 1
        else:
            self.eps.copy_(explicit["eps"])
            for cam in (cfg.cams if cfg.pixel else ()):
                self.masks_u8[cam].copy_(explicit["dropout"][cam])
        self.encode(st.params, feats_rows, state, self.Xp, self.F, self.masks_u8 if cfg.pixel else None, save=False,
                    save_proprio_actor=save and cfg.pixel)
        self.pol_state = state                                   # proprio input of the pass policy_backward differentiates
        self.policy_forward(st.params, self.Xp, save)
        ops.tanh_gaussian_fwd(self.mu, self.ls, self.eps, cfg.std_min, cfg.std_max, act_out, ld_act, self.logp, self.u, self.std, B, A)
        self.launches += 1

    def critic_loss_and_grads(self, keys, grad_scale=1.0, explicit=None):
        """sac.py:134-191 + its gradient w.r.t. group-0 parameters (written to store.grad)."""
        if self.fused is not None:
            return self.fused.critic_loss_and_grads(keys, grad_scale=grad_scale, explicit=explicit)
        cfg, B, E, st = self.cfg, self.B, self.cfg.ensemble, self.store
        obs_rows, next_rows = slice(0, B), slice(B, 2 * B)
        # three independent forward branches:
        #   side 0: Q(s, a) with params, saved for backward      side 1: target-encoder heads on s'
        #   main:   a', logp' ~ pi(s') (params, train=True), then Q'(s', a') with target params once side 1 has delivered enc(s')
        s0, s1 = self.side
        s0.fork()
        s1.fork()
        with s0:
            self.encode(st.params, obs_rows, self.state_o, self.Xc, self.FA, None, save=True, sc=self.sc_side[0])
            ops.copy2d(self.actions.data_ptr(), cfg.action_dim, ops.at(self.Xc, self.F), self.FA, B, cfg.action_dim)
            self.critic_forward(st.params, self.Xc, self.c_main, self.q, save=True, ws=self.ws_side[0])
        with s1:
            self.encode(st.target, next_rows, self.state_n, self.Xt, self.FA, None, save=False, sc=self.sc_side[1])
        self._policy_pass(next_rows, self.state_n, L.KEY_CRITIC_NEXT, L.KEY_CRITIC_NEXT, keys, ops.at(self.Xt, self.F), self.FA, save=False,
                          explicit=None if explicit is None else explicit["critic"])
        n_sub = 0
        if cfg.subsample is not None:
            if explicit is None:
                ops.subsample_idx(ops.key_ptr(keys, L.KEY_CRITIC_SUBSAMPLE), E, self.sub, cfg.subsample)
                self.launches += 1
            else:
                self.sub.copy_(explicit["critic"]["subsample"])
            n_sub = cfg.subsample
        s1.join()
        self.critic_forward(st.target, self.Xt, self.c_tgt, self.q_next, save=False)
        s0.join()
        ops.critic_loss(self.q, self.q_next, self.sub, n_sub, self.rewards, self.masks, self.logp, self.P(st.params, "modules_temperature/lagrange"),
                        cfg.backup_entropy, cfg.discount, grad_scale, self.target_q, self.dq, self.info.data_ptr(), E, B)
        self.critic_backward(self.Xc, self.c_main, self.dq, param_grads=True, need_dx=cfg.pixel)
        if cfg.pixel:
            self.encode_backward(self.dX, self.Xc, obs_rows, self.state_o)
        s0.join()                                               # weight / bias gradients
        self.launches += 2

    def actor_temp_loss_and_grads(self, keys, grad_scale=1.0, explicit=None, do_actor=True, do_temperature=True):
        """sac.py:193-234 + gradients w.r.t. group-1 / group-2 parameters (and the actor-tx twin of the proprio encoder)."""
        if self.fused is not None and os.environ.get("SERL_FUSED_ACTOR", "1") != "0":
            return self.fused.actor_temp_loss_and_grads(keys, grad_scale=grad_scale, explicit=explicit, do_actor=do_actor, do_temperature=do_temperature)
        cfg, B, E, A, st = self.cfg, self.B, self.cfg.ensemble, self.cfg.action_dim, self.store
        obs_rows, next_rows = slice(0, B), slice(B, 2 * B)
        lam = self.P(st.params, "modules_temperature/lagrange")
        if do_actor:
            self._actor_loss_and_grads(keys, grad_scale, explicit, obs_rows, lam)
        if do_temperature:
            # temperature: entropy of pi(.|s') with a fresh dropout mask / sample
            self._policy_pass(next_rows, self.state_n, L.KEY_TEMP_NEXT, L.KEY_TEMP_NEXT, keys, self.act_scratch.data_ptr(), A, save=False,
                              explicit=None if explicit is None else explicit["temperature"])
            ops.temperature_loss(self.logp, lam, cfg.target_entropy, grad_scale, self.P(st.grad, "modules_temperature/lagrange"), ops.at(self.info, 8), B)
            self.launches += 1

    def _actor_loss_and_grads(self, keys, grad_scale, explicit, obs_rows, lam):
        cfg, B, E, A, st = self.cfg, self.B, self.cfg.ensemble, self.cfg.action_dim, self.store
        # actor: a, logp ~ pi_theta(s); q = mean_e Q_e(s, a) with constant critic params
        self._policy_pass(obs_rows, self.state_o, L.KEY_ACTOR_SAMPLE, L.KEY_ACTOR_DROPOUT, keys, ops.at(self.Xc, self.F), self.FA, save=True,
                          explicit=None if explicit is None else explicit["actor"])
        self.encode(st.params, obs_rows, self.state_o, self.Xc, self.FA, None, save=False)
        self.critic_forward(st.params, self.Xc, self.c_main, self.q, save=True)
        ops.fill(self.dq.data_ptr(), -grad_scale / (E * B), E * B)
        self.critic_backward(self.Xc, self.c_main, self.dq, param_grads=False, need_dx=True)
        ops.actor_loss(self.q, self.logp, lam, ops.at(self.dX, self.F), self.FA, ops.at(self.Xc, self.F), self.FA, self.std, self.ls, self.eps,
                       cfg.std_min, cfg.std_max, grad_scale, self.dmu, self.dls, ops.at(self.info, 4), E, B, A)
        self.policy_backward(self.Xp)
        self.launches += 2

    def optimizer_step(self, live, polyak: bool):
        cfg, st = self.cfg, self.store
        ops.adam_polyak(st.params, st.target, st.m, st.v, st.grad, st.seg_end, live, st.counts, cfg.lr, cfg.warmup, cfg.tau, polyak,
                        lr_out=self.lr_info, n=st.n_main, gap=INFO_GAP, aux=(st.aux_lo, st.aux_hi, st.aux_off))
        self.launches += 2
