// Shared by the CUDA-core SGEMM (gemm_fp32.cu) and the tensor-core 3xTF32 GEMM (gemm_tf32x3.cu): kernel arguments and
// the deterministic split-K / batch-reduce pass (fixed summation order, no atomics).
#pragma once
#include "common.cuh"

namespace serl {

struct GemmArgs {
  const float* A; const float* B; float* C; const float* bias; float* ws;
  int M, N, K, Z, S;                 // S = k-splits
  long long sAz, sAm, sAk, sBz, sBk, sBn, sCz, sBiasZ;
  int ldc;
  int accumulate, to_ws;
  int debug, kchunk, a_mode, b_mode;                // tensor-core path only: staging mode of each operand (see gemm_tf32x3.cu)
};

// C[zc](m,n) = sum_{parts} ws[part](m,n) + bias + (accumulate ? C : 0); parts of zc: reduce_z ? all Z*S : S.
static __global__ void gemm_reduce_kernel(const GemmArgs g, int reduce_z) {
  pdl_prologue();
  const int ZC = reduce_z ? 1 : g.Z;
  const size_t MN = (size_t)g.M * g.N, total = MN * ZC;
  for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
    const int zc = (int)(e / MN); const size_t mn = e - (size_t)zc * MN;
    const int m = (int)(mn / g.N), n = (int)(mn - (size_t)m * g.N);
    const int p0 = reduce_z ? 0 : zc * g.S, np = reduce_z ? g.Z * g.S : g.S;
    float v = 0.f;
    for (int p = 0; p < np; ++p) v += g.ws[(size_t)(p0 + p) * MN + mn];
    if (g.bias) v += (g.bias + zc * g.sBiasZ)[n];
    float* c = g.C + zc * g.sCz + (size_t)m * g.ldc + n;
    *c = g.accumulate ? (*c + v) : v;
  }
}


// host side: launch the reduce pass for a descriptor whose partials sit in g.ws
inline int launch_gemm_reduce(const GemmArgs& g, int reduce_z, cudaStream_t st) {
  size_t total = (size_t)g.M * g.N * (reduce_z ? 1 : g.Z);
  int blocks = (int)((total + 255) / 256); if (blocks > 148 * 8) blocks = 148 * 8;
  launch_k(gemm_reduce_kernel, blocks, 256, 0, st, g, reduce_z);
  return check_launch("gemm_reduce_kernel");
}

}  // namespace serl
