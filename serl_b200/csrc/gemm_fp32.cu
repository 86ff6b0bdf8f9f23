// Batched, strided fp32 GEMM on CUDA cores with deterministic split-K / batch-reduce.
// Carries every dense contraction of the trainable heads in the fp32 (1e-5 parity) build:
//   flax nn.Dense forward / input-grad / weight-grad for the encoder bottleneck (4096->256), proprio
//   (S->64), the vmapped critic ensemble (E x [F+A -> 256 -> 256]) and the policy MLP.
// Reference call sites (relative to serl_launcher/serl_launcher): networks/mlp.py:22-31,
// networks/actor_critic_nets.py:64-72,187-192, vision/resnet_v1.py:371, common/encoding.py:65-67;
// backward = what jax.grad emits for those (common/common.py:204).
//
//   C[z](m, n) = sum_k A[z](m, k) * B[z](k, n) (+ bias[z](n)) (+ C[z](m, n) if accumulate)
// with element addresses A + z*sAz + m*sAm + k*sAk, B + z*sBz + k*sBk + n*sBn, C + z*sCz + m*ldc + n.
// reduce_z: a single C = sum_z (...)  (used for dX = sum_e dZ_e W_e^T of the broadcast ensemble input).
#include "gemm_common.cuh"
#include "serl_b200.h"

namespace serl {

constexpr int GM = 64, GN = 64, GK = 16;

constexpr int GSTAGES = 4;

// 4-stage cp.async pipeline: these GEMMs are small (a few hundred CTAs of 64x64 tiles), so each CTA is bound by the
// global->shared latency of its k-loop; with three k-blocks of loads in flight per CTA the loop runs at FMA speed.
__global__ void __launch_bounds__(256) gemm_f32_kernel(const GemmArgs g) {
  pdl_prologue();
  __shared__ __align__(16) float As[GSTAGES][GK][GM + 4];
  __shared__ __align__(16) float Bs[GSTAGES][GK][GN + 4];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int z = blockIdx.z / g.S, s = blockIdx.z - z * g.S;
  const int m0 = blockIdx.x * GM, n0 = blockIdx.y * GN;
  const int kchunk = ceil_div(ceil_div(g.K, g.S), GK) * GK;
  const int kbeg = s * kchunk, kend = min(g.K, kbeg + kchunk);
  const float* A = g.A + z * g.sAz;
  const float* B = g.B + z * g.sBz;

  // thread -> tile element mappings, chosen so the unit-stride dimension is fastest across threads
  const bool a_kfast = (g.sAk == 1);
  const bool b_nfast = (g.sBn == 1);
  auto issue = [&](int k0, int buf) {
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      int m, k;
      if (a_kfast) { k = tid & 15; m = (tid >> 4) + 16 * e; } else { m = tid & 63; k = (tid >> 6) + 4 * e; }
      const int gm = m0 + m, gk = k0 + k;
      const bool va = gm < g.M && gk < kend;
      const float* pa = va ? A + gm * g.sAm + gk * g.sAk : g.A;
      asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"((uint32_t)__cvta_generic_to_shared(&As[buf][k][m])), "l"(pa), "r"(va ? 4u : 0u) : "memory");
      int n, kb;
      if (b_nfast) { n = tid & 63; kb = (tid >> 6) + 4 * e; } else { kb = tid & 15; n = (tid >> 4) + 16 * e; }
      const int gn = n0 + n, gkb = k0 + kb;
      const bool vb = gn < g.N && gkb < kend;
      const float* pb = vb ? B + gkb * g.sBk + gn * g.sBn : g.B;
      asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"((uint32_t)__cvta_generic_to_shared(&Bs[buf][kb][n])), "l"(pb), "r"(vb ? 4u : 0u) : "memory");
    }
  };

  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  const int nk = kend > kbeg ? ceil_div(kend - kbeg, GK) : 0;
#pragma unroll
  for (int st = 0; st < GSTAGES - 1; ++st) {
    if (st < nk) issue(kbeg + st * GK, st);
    asm volatile("cp.async.commit_group;" ::: "memory");
  }
  for (int kt = 0; kt < nk; ++kt) {
    asm volatile("cp.async.wait_group %0;" ::"n"(GSTAGES - 2) : "memory");
    __syncthreads();                                        // stage kt landed for everyone; stage (kt-1) fully consumed
    if (kt + GSTAGES - 1 < nk) issue(kbeg + (kt + GSTAGES - 1) * GK, (kt + GSTAGES - 1) % GSTAGES);
    asm volatile("cp.async.commit_group;" ::: "memory");
    const int buf = kt % GSTAGES;
#pragma unroll
    for (int k = 0; k < GK; ++k) {
      float4 av = *reinterpret_cast<const float4*>(&As[buf][k][ty * 4]);
      float4 bv = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 4]);
      const float aa[4] = {av.x, av.y, av.z, av.w}, bb[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(aa[i], bb[j], acc[i][j]);
    }
  }
  asm volatile("cp.async.wait_all;" ::: "memory");

  if (g.to_ws) {
    float* P = g.ws + ((size_t)blockIdx.z * g.M) * g.N;        // part index = z*S + s
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int m = m0 + ty * 4 + i; if (m >= g.M) continue;
#pragma unroll
      for (int j = 0; j < 4; ++j) { const int n = n0 + tx * 4 + j; if (n < g.N) P[(size_t)m * g.N + n] = acc[i][j]; }
    }
  } else {
    float* C = g.C + z * g.sCz;
    const float* bias = g.bias ? g.bias + z * g.sBiasZ : nullptr;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int m = m0 + ty * 4 + i; if (m >= g.M) continue;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int n = n0 + tx * 4 + j; if (n >= g.N) continue;
        float v = acc[i][j] + (bias ? bias[n] : 0.f);
        float* c = C + (size_t)m * g.ldc + n;
        *c = g.accumulate ? (*c + v) : v;
      }
    }
  }
}

}  // namespace serl

using namespace serl;

extern "C" int serl_gemm_f32(const serl_gemm_desc* d, void* stream) {
  if (!d || d->M < 1 || d->N < 1 || d->K < 1 || d->Z < 1 || !d->A || !d->B || !d->C) {
    set_last_error("serl_gemm_f32: invalid descriptor");
    return SERL_ERR_INVALID;
  }
  GemmArgs g{};
  g.A = d->A; g.B = d->B; g.C = d->C; g.bias = d->bias; g.ws = d->workspace;
  g.M = d->M; g.N = d->N; g.K = d->K; g.Z = d->Z;
  g.sAz = d->sAz; g.sAm = d->sAm; g.sAk = d->sAk; g.sBz = d->sBz; g.sBk = d->sBk; g.sBn = d->sBn;
  g.sCz = d->sCz; g.sBiasZ = d->sBiasZ; g.ldc = d->ldc; g.accumulate = d->accumulate;
  const int tiles = ceil_div(d->M, GM) * ceil_div(d->N, GN) * d->Z;
  // every CTA is latency-bound, so aim for ~4 CTAs per SM: split K (deterministic partial sums) until ~600 CTAs exist
  int S = 1;
  if (tiles < 448 && d->K >= 128) {
    S = ceil_div(600, tiles);
    if (S > d->K / 64) S = d->K / 64;
    if (S < 1) S = 1;
  }
  const size_t part = (size_t)d->M * d->N * sizeof(float);
  if (d->reduce_z || S > 1) {
    while (S > 1 && part * (size_t)d->Z * S > d->workspace_bytes) --S;
    if ((d->reduce_z || S > 1) && (!d->workspace || part * (size_t)d->Z * S > d->workspace_bytes)) {
      if (d->reduce_z) { set_last_error("serl_gemm_f32: reduce_z needs %zu workspace bytes", part * (size_t)d->Z); return SERL_ERR_INVALID; }
      S = 1;
    }
  }
  g.S = S;
  g.to_ws = (d->reduce_z || S > 1) ? 1 : 0;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  dim3 grid(ceil_div(d->M, GM), ceil_div(d->N, GN), d->Z * S);
  launch_k(gemm_f32_kernel, grid, 256, 0, st, g);
  if (int e = check_launch("gemm_f32_kernel")) return e;
  if (g.to_ws) {
    return launch_gemm_reduce(g, d->reduce_z, st);
  }
  return SERL_OK;
}
