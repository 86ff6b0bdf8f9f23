// Stride-1 3x3 convolution on tcgen05 tensor cores WITHOUT im2col redundancy ("shifted window").
//
// The (zero-padded) input of an image is viewed as a 1-D raster with pitch P = W+1 and R = H+1 rows per image: the
// extra column / row are zeros shared between neighbours (column W of row h is both the right pad of row h and the left
// pad of row h+1; row H of image n is both its bottom pad and the top pad of image n+1).  In that raster the input of
// output position q under tap (r, s) is simply position q + (r-1)*P + (s-1), so for a tile of 128 consecutive raster
// positions ALL NINE taps read windows of ONE contiguous patch of 128 + 2P + 2 positions.  The producers stage that
// patch once per 64-channel block into 128B-swizzled shared memory (one 128-byte row per position) and the nine taps are
// nine UMMA descriptors whose start address is shifted by (r*P + s) rows - no data is gathered twice.  Weights arrive by
// TMA (cp.async.bulk.tensor, 128B swizzle) through their own mbarrier ring.  Outputs at pad positions are computed and
// discarded (M efficiency H*W / ((H+1)(W+1)): 94% at 32x32, 64% at 4x4).
//
//   warps 0-3  epilogue (tcgen05.ld, GroupNorm partial sums segmented by image, 16-bit pack, NHWC store of valid positions)
//   warps 4-7  stage the patch with cp.async (coalesced 16-byte copies, zero-fill at pad positions)
//   warp 8     tcgen05.mma issuer (9 taps x 4 K-steps per channel block), commits
//   warp 9     TMA issuer for the weight tiles
//
// Replaces the conv of vision/resnet_v1.py:142-147 (ResNetBlock 3x3 convs with stride 1, SAME padding).
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "common.cuh"
#include "serl_b200.h"

namespace serl {

struct Conv3Args {
  const uint16_t* x;       // (N,H,W,Ci) 16-bit, already activated
  uint16_t* y;             // (N,H,W,Co) raw conv output
  float* stats;            // (N,4,2)
  int32_t* error;
  int N, H, W, Ci, Co, P, R, cblocks, Cg, Lp, patch_bytes, base_offset_mode;
  long long Q;
};

struct C3Bf16 {
  static constexpr uint32_t kUmmaFormat = 1;
  __device__ static inline uint32_t pack(float lo, float hi) { __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi); return *reinterpret_cast<uint32_t*>(&v); }
};
struct C3Fp16 {
  static constexpr uint32_t kUmmaFormat = 0;
  __device__ static inline uint32_t pack(float lo, float hi) {
    __half2 v = __floats2half2_rn(fminf(fmaxf(lo, -65504.f), 65504.f), fminf(fmaxf(hi, -65504.f), 65504.f));
    return *reinterpret_cast<uint32_t*>(&v);
  }
};

__device__ inline uint32_t c3_smem(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ inline void c3_mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(c3_smem(bar)), "r"(count));
}
__device__ inline void c3_mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(c3_smem(bar)) : "memory");
}
__device__ inline void c3_mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(c3_smem(bar)), "r"(bytes) : "memory");
}
__device__ inline bool c3_mbar_wait(uint64_t* bar, uint32_t parity, int32_t* error) {
  const uint32_t addr = c3_smem(bar);
  const long long t0 = clock64();
#pragma unroll 1
  for (;;) {
    uint32_t done;
    asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                 : "=r"(done) : "r"(addr), "r"(parity) : "memory");
    if (done) return true;
    if (clock64() - t0 > 4000000000ll) break;
  }
  atomicOr(error, 4);
  return false;
}
__device__ inline uint64_t c3_desc(uint32_t saddr, uint32_t base_offset) {
  return (uint64_t)((saddr & 0x3FFFF) >> 4) | (1ull << 16) | (64ull << 32) | (1ull << 46) | ((uint64_t)(base_offset & 7) << 49) | (2ull << 61);
}
__device__ inline void c3_mma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile("{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}"
               ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ inline void c3_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(c3_smem(bar)) : "memory");
}
__device__ inline void c3_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                 "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
               : "r"(taddr) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ inline void c3_tma_2d(void* smem_dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               ::"r"(c3_smem(smem_dst)), "l"(map), "r"(c3_smem(bar)), "r"(c0), "r"(c1) : "memory");
}

constexpr int C3_MAXR = 13;           // patch rows per producer thread: ceil((128 + 2*33 + 2) / 16)
constexpr int C3_THREADS = 320;       // warps 0-3 epilogue | 4-7 patch producers | 8 MMA issuer | 9 weight-TMA issuer

// Persistent: each CTA walks tiles blockIdx.x, +gridDim.x, ...; the four roles run decoupled through mbarrier rings
// (patch full/empty x2, weight full/empty xBSTAGES, accumulator full/empty x2 - TMEM holds two accumulators), so the
// gather of tile i+1, the MMAs of tile i and the epilogue of tile i-1 overlap.
// kResidentB (Ci == 64, Co == BN): the whole 9-tap weight tensor (9 x BN x 128 B) is loaded ONCE per CTA and stays in
// shared memory - a persistent CTA then streams only input patches (1 CTA / SM, BSTAGES must be 9).
//
// kCoalEpi (EXPERIMENTAL, opt-in with SERL_EPI_COAL=1, not yet validated on hardware - see DESIGN.md section 8): the default
// epilogue lets every thread store its own output row, so one STG.128 touches 32 different 128-byte lines = 32 L1TEX
// data-pipe wavefronts; ncu shows those stores taking 29 % of the data pipe that also feeds the tensor cores' operand
// reads (49 %) and the weight TMA writes.  The variant transposes 64-channel groups through a per-warp 4 KB shared tile
// so that a store instruction covers 4 rows x 128 contiguous bytes (4 wavefronts): 96 instead of 256 wavefronts per group.
constexpr int C3_EPI_STAGE = 4 * 32 * 128;             // 4 epilogue warps x 32 rows x 128 B

template <class F, int BN, int BSTAGES, int PSTAGES, bool kResidentB, bool kCoalEpi = false>
__global__ void __launch_bounds__(C3_THREADS, kResidentB ? 1 : 2) conv3x3_tc_kernel(const __grid_constant__ CUtensorMap wmap, const Conv3Args a) {
  pdl_prologue();
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  constexpr int B_STAGE = BN * 128;
  uint8_t* sP = smem;                                   // PSTAGES patches
  uint8_t* sB = smem + PSTAGES * a.patch_bytes;
  uint8_t* sEpi = sB + BSTAGES * B_STAGE;               // kCoalEpi: per-warp output transpose tiles
  uint64_t* pfull = reinterpret_cast<uint64_t*>(sEpi + (kCoalEpi ? C3_EPI_STAGE : 0));
  uint64_t* pempty = pfull + PSTAGES;
  uint64_t* bfull = pempty + PSTAGES;
  uint64_t* bempty = bfull + BSTAGES;
  uint64_t* afull = bempty + BSTAGES;
  uint64_t* aempty = afull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(aempty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_tiles_n = a.Co / BN;
  const int n_tiles = (int)((a.Q + 127) / 128) * n_tiles_n;

  if (threadIdx.x == 0) {
    for (int s = 0; s < PSTAGES; ++s) { c3_mbar_init(&pfull[s], 128); c3_mbar_init(&pempty[s], 1); }
    for (int s = 0; s < 2; ++s) { c3_mbar_init(&afull[s], 1); c3_mbar_init(&aempty[s], 4); }
    for (int s = 0; s < BSTAGES; ++s) { c3_mbar_init(&bfull[s], 1); c3_mbar_init(&bempty[s], 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 8) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(c3_smem(tmem_slot)), "r"((uint32_t)(2 * BN)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (warp == 9 && lane == 0) asm volatile("prefetch.tensormap [%0];" ::"l"(&wmap) : "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;
  const int per_img = a.R * a.P;

  if (warp >= 4 && warp < 8) {
    // ------------------------------- patch producers -------------------------------
    const int tid = threadIdx.x - 128, chunk = tid & 7, rsub = tid >> 3;
    bool ok = true;
    int pc = 0;                                            // patches produced so far (ring position)
    for (int tile = blockIdx.x; tile < n_tiles && ok; tile += gridDim.x) {
      const int q0 = (tile / n_tiles_n) * 128;
      int poff[C3_MAXR];                                   // element offset of the pixel (channel 0) or -1 for a zero row
      {
        int q = q0 - a.P - 1 + rsub + per_img;             // shifted by one image so it is non-negative
        int n = q / per_img - 1; const int rem = q - (n + 1) * per_img;
        int hrow = rem / a.P; int wcol = rem - hrow * a.P;
#pragma unroll
        for (int i = 0; i < C3_MAXR; ++i) {                // walk the raster 16 positions at a time
          const int j = rsub + 16 * i;
          poff[i] = -1;
          if (j < a.Lp && n >= 0 && n < a.N && wcol < a.W && hrow < a.H) poff[i] = ((n * a.H + hrow) * a.W + wcol) * a.Ci;
          wcol += 16;
          while (wcol >= a.P) { wcol -= a.P; ++hrow; }
          while (hrow >= a.R) { hrow -= a.R; ++n; }
        }
      }
      for (int cb = 0; cb < a.cblocks && ok; ++cb, ++pc) {
        const int ps = pc % PSTAGES;
        ok = c3_mbar_wait(&pempty[ps], (uint32_t)((pc / PSTAGES) & 1) ^ 1u, a.error);
        const uint32_t Ps = c3_smem(sP + ps * a.patch_bytes);
#pragma unroll
        for (int i = 0; i < C3_MAXR; ++i) {                 // cp.async: no register staging, zero-fill at pad positions
          const int j = rsub + 16 * i;
          if (j < a.Lp) {
            const bool inb = poff[i] >= 0;
            const uint16_t* src = a.x + (inb ? (size_t)(uint32_t)(poff[i] + cb * 64 + chunk * 8) : 0);
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;"
                         ::"r"(Ps + (j >> 3) * 1024 + (j & 7) * 128 + ((chunk ^ (j & 7)) << 4)), "l"(src), "r"(inb ? 16u : 0u) : "memory");
          }
        }
        asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(c3_smem(&pfull[ps])) : "memory");
      }
    }
    asm volatile("cp.async.wait_all;" ::: "memory");
  } else if (warp < 4) {
    // ------------------------------- epilogue ---------------------------------------
    bool ok = true;
    int ac = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++ac) {
      const int as = ac & 1;
      const int q0 = (tile / n_tiles_n) * 128, n0 = (tile % n_tiles_n) * BN;
      ok = ok && c3_mbar_wait(&afull[as], (uint32_t)((ac >> 1) & 1), a.error);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const int row = warp * 32 + lane;
      const int q = q0 + row;
      bool valid = false; int n_img = a.N - 1; int opix = 0;
      if (q < (int)a.Q) {
        n_img = q / per_img; const int rem = q - n_img * per_img;
        const int hrow = rem / a.P, wcol = rem - hrow * a.P;
        valid = wcol < a.W && hrow < a.H;
        opix = (n_img * a.H + hrow) * a.W + wcol;
      }
      valid = valid && ok;
      constexpr int MAXG = 4;
      float gs[MAXG], gss[MAXG];
#pragma unroll
      for (int g = 0; g < MAXG; ++g) { gs[g] = 0.f; gss[g] = 0.f; }
      const int g_first = n0 / a.Cg;
#pragma unroll
      for (int c0 = 0; c0 < BN; c0 += 16) {
        uint32_t v[16];
        c3_ld16(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(as * BN + c0), v);
        float s = 0.f, ss = 0.f;
        uint32_t pk[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float f0 = __uint_as_float(v[2 * j]), f1 = __uint_as_float(v[2 * j + 1]);
          s += f0 + f1; ss += f0 * f0 + f1 * f1;
          pk[j] = F::pack(f0, f1);
        }
        if (valid) {
          const int g = (n0 + c0) / a.Cg - g_first;
#pragma unroll
          for (int gg = 0; gg < MAXG; ++gg) if (gg == g) { gs[gg] += s; gss[gg] += ss; }
          if constexpr (!kCoalEpi) {
            uint4* dst = reinterpret_cast<uint4*>(a.y + (size_t)opix * a.Co + n0 + c0);
            dst[0] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
            dst[1] = make_uint4(pk[4], pk[5], pk[6], pk[7]);
          }
        }
        if constexpr (kCoalEpi) {
          // row = lane; 16-byte chunk ch of the 64-channel group at lane*128 + ((ch ^ (lane & 7)) << 4): conflict-free both ways
          uint8_t* tile = sEpi + warp * (32 * 128);
          const int ch = (c0 & 48) >> 3;
          *reinterpret_cast<uint4*>(tile + lane * 128 + ((ch ^ (lane & 7)) << 4)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
          *reinterpret_cast<uint4*>(tile + lane * 128 + (((ch + 1) ^ (lane & 7)) << 4)) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
          if ((c0 & 48) == 48) {                                 // 64 channels staged: store them 4 rows x 128 B per instruction
            __syncwarp();
            const int cgrp = c0 - 48;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const int r = (lane >> 3) + 4 * i, cc = lane & 7;
              const int opix_r = __shfl_sync(0xffffffffu, opix, r);
              const int valid_r = __shfl_sync(0xffffffffu, (int)valid, r);
              const uint4 v4 = *reinterpret_cast<const uint4*>(tile + r * 128 + ((cc ^ (r & 7)) << 4));
              if (valid_r) *reinterpret_cast<uint4*>(a.y + (size_t)opix_r * a.Co + n0 + cgrp + cc * 8) = v4;
            }
            __syncwarp();
          }
        }
      }
      // accumulator drained: hand the TMEM stage back before the (slower) statistics
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) c3_mbar_arrive(&aempty[as]);
      // GroupNorm sums, segmented by image: a warp's 32 consecutive raster positions touch at most 3 images
      const int ngroups = (BN + a.Cg - 1) / a.Cg;
      const int id_lo = __shfl_sync(0xffffffffu, n_img, 0), id_hi = __shfl_sync(0xffffffffu, n_img, 31);
      for (int id = id_lo; id <= id_hi; ++id) {
        for (int g = 0; g < ngroups; ++g) {
          float s = 0.f, ss = 0.f;
#pragma unroll
          for (int gg = 0; gg < MAXG; ++gg) if (gg == g && n_img == id && valid) { s = gs[gg]; ss = gss[gg]; }
          s = warp_sum(s); ss = warp_sum(ss);
          if (lane == 0 && (s != 0.f || ss != 0.f)) {
            float* st = a.stats + ((size_t)id * 4 + g_first + g) * 2;
            atomicAdd(st, s); atomicAdd(st + 1, ss);
          }
        }
      }
    }
  } else if (warp == 8) {
    // ------------------------------- MMA issuer -----------------------------------
    // ONE thread runs the whole issue loop (no per-iteration warp reconvergence): the loop body is a handful of integer
    // ops per tcgen05.mma, which matters because a 128x64x16 MMA only occupies the tensor pipe for ~32 cycles.
    if (lane == 0) {
      const uint32_t idesc = (1u << 4) | (F::kUmmaFormat << 7) | (F::kUmmaFormat << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      const uint64_t desc_hi = (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);     // LBO=1, SBO=1024 B, version 1, SWIZZLE_128B
      const uint32_t p_lo = (c3_smem(sP) & 0x3FFFF) >> 4, b_lo = (c3_smem(sB) & 0x3FFFF) >> 4;  // start addresses in 16-byte units
      const uint32_t p_step = (uint32_t)a.patch_bytes >> 4, b_step = (uint32_t)B_STAGE >> 4;
      uint32_t tap_lo[9];
#pragma unroll
      for (int t = 0; t < 9; ++t) tap_lo[t] = (uint32_t)((t / 3) * a.P + (t % 3)) * 8u;        // window shift: rows x 128 B / 16
      bool ok = true;
      int it = 0, pc = 0, ac = 0;
      for (int tile = blockIdx.x; tile < n_tiles && ok; tile += gridDim.x, ++ac) {
        const int as = ac & 1;
        ok = c3_mbar_wait(&aempty[as], (uint32_t)((ac >> 1) & 1) ^ 1u, a.error);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t tmem_d = tmem_base + (uint32_t)(as * BN);
        for (int cb = 0; cb < a.cblocks && ok; ++cb, ++pc) {
          const int ps = pc % PSTAGES;
          ok = c3_mbar_wait(&pfull[ps], (uint32_t)((pc / PSTAGES) & 1), a.error);
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");    // cp.async (generic proxy) writes -> tensor-core reads
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t pbase = p_lo + (uint32_t)ps * p_step;
#pragma unroll
          for (int tap = 0; tap < 9; ++tap, ++it) {
            const int sb = kResidentB ? tap : it % BSTAGES;
            if (!kResidentB || it < 9) {                              // resident weights: each tap's tile is waited for once
              ok = ok && c3_mbar_wait(&bfull[sb], kResidentB ? 0u : (uint32_t)((it / BSTAGES) & 1), a.error);
              asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            }
            const uint64_t ad = desc_hi | (uint64_t)(pbase + tap_lo[tap]);
            const uint64_t bd = desc_hi | (uint64_t)(b_lo + (uint32_t)sb * b_step);
#pragma unroll
            for (int k = 0; k < 4; ++k) c3_mma(tmem_d, ad + 2 * k, bd + 2 * k, idesc, (uint32_t)((cb | tap | k) != 0));
            if (!kResidentB) c3_commit(&bempty[sb]);
          }
          c3_commit(&pempty[ps]);
        }
        if (ok) c3_commit(&afull[as]); else c3_mbar_arrive(&afull[as]);
      }
    }
  } else {
    // ------------------------------- weight TMA issuer ----------------------------
    if (lane == 0) {
      if (kResidentB) {
        for (int tap = 0; tap < 9; ++tap) {                         // one-time load of the whole weight tensor
          c3_mbar_expect_tx(&bfull[tap], (uint32_t)B_STAGE);
          c3_tma_2d(sB + tap * B_STAGE, &wmap, tap * a.Ci, 0, &bfull[tap]);
        }
      } else {
        bool ok = true;
        int it = 0;
        for (int tile = blockIdx.x; tile < n_tiles && ok; tile += gridDim.x) {
          const int n0 = (tile % n_tiles_n) * BN;
          for (int cb = 0; cb < a.cblocks && ok; ++cb) {
            for (int tap = 0; tap < 9 && ok; ++tap, ++it) {
              const int sb = it % BSTAGES;
              ok = c3_mbar_wait(&bempty[sb], (uint32_t)((it / BSTAGES) & 1) ^ 1u, a.error);
              if (!ok) break;
              c3_mbar_expect_tx(&bfull[sb], (uint32_t)B_STAGE);
              c3_tma_2d(sB + sb * B_STAGE, &wmap, tap * a.Ci + cb * 64, n0, &bfull[sb]);
            }
          }
        }
      }
    }
  }
  __syncthreads();
  if (warp == 8) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)(2 * BN)) : "memory");
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess && qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

template <class F, int BN, int BSTAGES, int PSTAGES, bool kResidentB, bool kCoalEpi = false>
static int launch_conv3(const CUtensorMap& map, const Conv3Args& a, cudaStream_t st) {
  const size_t smem = (size_t)PSTAGES * a.patch_bytes + (size_t)BSTAGES * BN * 128 + (kCoalEpi ? C3_EPI_STAGE : 0) + 1024 + 256;
  auto kern = conv3x3_tc_kernel<F, BN, BSTAGES, PSTAGES, kResidentB, kCoalEpi>;
  static size_t configured = 0;
  if (smem > configured) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return check_launch("cudaFuncSetAttribute(conv3x3_tc)");
    configured = smem;
  }
  static int sms = 0;
  if (!sms) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev); if (sms <= 0) sms = 148; }
  const long long tiles = ((a.Q + 127) / 128) * (a.Co / BN);
  const long long slots = (kResidentB ? 1ll : 2ll) * sms;               // persistent CTAs per device
  const int grid = (int)(tiles < slots ? tiles : slots);
  launch_k(kern, grid, C3_THREADS, smem, st, map, a);
  return check_launch("conv3x3_tc_kernel");
}

}  // namespace serl

using namespace serl;

extern "C" int serl_conv3x3s1_tc_h16(const serl_conv_tc_desc* d, int base_offset_mode, void* stream) {
  if (!d || !d->x || !d->w || !d->y || !d->stats || !d->error) { set_last_error("serl_conv3x3s1_tc_h16: invalid descriptor"); return SERL_ERR_INVALID; }
  if (d->kh != 3 || d->kw != 3 || d->stride != 1 || d->pad_lo != 1 || d->Ho != d->Hi || d->Wo != d->Wi || d->Ci % 64 || d->Co % 64 || d->Wi > 32 || d->in_a) {
    set_last_error("serl_conv3x3s1_tc_h16: needs a 3x3 stride-1 SAME conv, Ci,Co %% 64 == 0, W <= 32, no operand transform"); return SERL_ERR_UNSUPPORTED;
  }
  EncodeTiledFn enc = get_encode();
  if (!enc) { set_last_error("serl_conv3x3s1_tc_h16: cuTensorMapEncodeTiled unavailable"); return SERL_ERR_CUDA; }
  Conv3Args a{};
  a.x = static_cast<const uint16_t*>(d->x); a.y = static_cast<uint16_t*>(d->y); a.stats = d->stats; a.error = d->error;
  a.N = d->N; a.H = d->Hi; a.W = d->Wi; a.Ci = d->Ci; a.Co = d->Co; a.P = d->Wi + 1; a.R = d->Hi + 1;
  a.cblocks = d->Ci / 64; a.Cg = d->Co / 4; a.Lp = 128 + 2 * a.P + 2; a.patch_bytes = ((a.Lp * 128 + 1023) / 1024) * 1024;
  a.Q = (long long)d->N * a.R * a.P; a.base_offset_mode = base_offset_mode;
  if (a.Q + 4096 >= (1ll << 31) || (long long)d->N * d->Hi * d->Wi * d->Ci >= (1ll << 31)) { set_last_error("serl_conv3x3s1_tc_h16: tensor too large for 32-bit raster indexing"); return SERL_ERR_UNSUPPORTED; }
  const int BN = d->Co == 64 ? 64 : 128;
  CUtensorMap map;
  const cuuint64_t gdim[2] = {(cuuint64_t)9 * d->Ci, (cuuint64_t)d->Co};
  const cuuint64_t gstr[1] = {(cuuint64_t)9 * d->Ci * 2};
  const cuuint32_t box[2] = {64u, (cuuint32_t)BN};
  const cuuint32_t estr[2] = {1u, 1u};
  CUresult r = enc(&map, d->fmt == SERL_FMT_FP16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(d->w),
                   gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_last_error("serl_conv3x3s1_tc_h16: cuTensorMapEncodeTiled failed (%d)", (int)r); return SERL_ERR_CUDA; }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  // The weight ring is what bounds these kernels (TMA latency x ring depth vs 9 tiles per channel block), so it gets the
  // shared memory: BN=64: 2 patches (25 KiB) + 7 weight stages (8 KiB); BN=128: 2 patches (<=21 KiB) + 4 weight stages (16 KiB).
  //            with Ci == 64 as well the 72 KiB weight tensor stays resident (1 CTA / SM, 4 patches)
  const bool resident = false && BN == 64 && d->Ci == 64;   // measured slower than two streaming CTAs per SM (profiles/r01_trunk_kernels.md)
  static int coal = -1;                                   // EXPERIMENTAL coalesced epilogue, off unless SERL_EPI_COAL=1
  if (coal < 0) { const char* e = getenv("SERL_EPI_COAL"); coal = (e && atoi(e) != 0) ? 1 : 0; }
  if (coal && d->fmt == SERL_FMT_FP16)                     // 16 KB of the weight ring goes to the transpose tiles
    return BN == 64 ? launch_conv3<C3Fp16, 64, 5, 2, false, true>(map, a, st) : launch_conv3<C3Fp16, 128, 3, 2, false, true>(map, a, st);
  if (d->fmt == SERL_FMT_FP16) {
    if (resident) return launch_conv3<C3Fp16, 64, 9, 4, true>(map, a, st);
    return BN == 64 ? launch_conv3<C3Fp16, 64, 7, 2, false>(map, a, st) : launch_conv3<C3Fp16, 128, 4, 2, false>(map, a, st);
  }
  if (resident) return launch_conv3<C3Bf16, 64, 9, 4, true>(map, a, st);
  return BN == 64 ? launch_conv3<C3Bf16, 64, 7, 2, false>(map, a, st) : launch_conv3<C3Bf16, 128, 4, 2, false>(map, a, st);
}
