"""Times serl_tgemm_tf32 against serl_gemm_tf32x3 on the heads' shapes (20 launches per CUDA graph, CUDA events)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from serl_b200 import _lib as L, ops


def timeit(fn, reps=20):
    fn(); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps):
            fn()
    g.replay(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5):
        g.replay()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / (5 * reps) * 1e3


ws3, wst = ops.Workspace(64 << 20, "cuda", "tf32x3"), ops.Workspace(64 << 20, "cuda")
for name, M, K, N, Z, bcast in [("critic L1 fwd", 256, 580, 256, 10, True), ("critic L2 fwd", 256, 256, 256, 10, False),
                                ("enc dense fwd", 256, 4096, 256, 1, False), ("policy L1", 256, 576, 256, 1, False),
                                ("critic L1 fwd B=32", 32, 580, 256, 10, True)]:
    x = torch.randn(1 if bcast else Z, M, K, device="cuda"); w = torch.randn(Z, K, N, device="cuda") / K ** 0.5; b = torch.randn(Z, N, device="cuda")
    out = torch.empty(Z, M, N, device="cuda")
    xz = 0 if bcast else M * K
    t3 = timeit(lambda: ops.dense_fwd(ws3, x.data_ptr(), K, w.data_ptr(), b.data_ptr(), out.data_ptr(), N, M, K, N, Z=Z, x_z=xz, out_z=M * N))
    p = ops.tgemm_problem(x.data_ptr(), w.data_ptr(), sAm=K, sAk=1, sBk=N, sBn=1, Z=Z, sAz=xz, sBz=K * N, C_=out.data_ptr(), sCz=M * N, ldc=N,
                          bias=b.data_ptr(), sBiasZ=N)
    tt = timeit(lambda: ops.tgemm(wst, [p], M, N, K))
    line = f"{name:22s} M={M} K={K} N={N} Z={Z}: tf32x3 {t3:6.1f} us   tgemm {tt:6.1f} us"
    if K <= 1024:
        sc = torch.ones(Z, N, device="cuda"); lb = torch.zeros(Z, N, device="cuda"); xh = torch.empty(Z, M, N, device="cuda"); rs = torch.empty(Z, M, device="cuda")
        p2 = ops.tgemm_problem(x.data_ptr(), w.data_ptr(), sAm=K, sAk=1, sBk=N, sBn=1, Z=Z, sAz=xz, sBz=K * N, C_=out.data_ptr(), sCz=M * N, ldc=N,
                               bias=b.data_ptr(), sBiasZ=N, ln_scale=sc.data_ptr(), ln_bias=lb.data_ptr(), sLnZ=N, xhat=xh.data_ptr(), rstd=rs.data_ptr(),
                               sXhatZ=M * N, sRstdZ=M)
        tl = timeit(lambda: ops.tgemm(None, [p2], M, N, K, epilogue=L.TGEMM_LN_TANH))
        z = torch.empty(Z * M, N, device="cuda")
        tl3 = timeit(lambda: (ops.dense_fwd(ws3, x.data_ptr(), K, w.data_ptr(), b.data_ptr(), z.data_ptr(), N, M, K, N, Z=Z, x_z=xz, out_z=M * N),
                              ops.ln_tanh_fwd(z.data_ptr(), N, sc.data_ptr(), lb.data_ptr(), M, N, out.data_ptr(), N, xh.data_ptr(), rs.data_ptr(), Z * M, N)))
        line += f"   | + LN + tanh: separate {tl3:6.1f} us   fused {tl:6.1f} us"
    print(line, flush=True)
