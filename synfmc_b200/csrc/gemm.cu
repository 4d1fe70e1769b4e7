// Linear-layer GEMM for the FMC denoising step:  C[M,N] = epilogue(A[M,K] * W[N,K]^T)
//
// This one kernel carries every `nn.Linear` / 1x1-conv on the hot path once Domain-LoRA has been
// folded into W (reference: fmc/models/attention_processor.py:138-157 q/k/v/out projections,
// fmc/models/motion_module.py:219,228 proj_in/proj_out, diffusers FeedForward GEGLU,
// PoseAdaptorAttnProcessor.qkv_merge attention_processor.py:257).
//
// sm_100a design: persistent CTAs (one per SM), warp-specialised
//   warp 0      TMA producer   (cp.async.bulk.tensor, SWIZZLE_128B, BK = 64 bf16 per stage)
//   warp 1      MMA issuer     (tcgen05.mma cta_group::1 kind::f16, 128 x BN x 16 per instruction)
//   warp 2      TMEM allocator
//   warps 4-7   epilogue       (tcgen05.ld -> bias / GEGLU / row-bias / residual -> global)
// The fp32 accumulator is double-buffered in TMEM so the epilogue of tile i overlaps the MMAs of tile i+1.
#include <cuda_bf16.h>
#include <cstdlib>

#include "common.cuh"
#include "ptx.cuh"

namespace fmc {

constexpr int GEMM_BM = 128;
constexpr int GEMM_BK = 64;
constexpr int GEMM_THREADS = 256;

struct GemmParams {
  int M, N, K;
  void* C;
  long long ldc;
  const float* bias;
  const __nv_bfloat16* residual;
  long long ldr;
  const float* rowbias;
  int rows_per_group;
  long long ldrb;
  int flags;
  int f16_col0;  // output columns >= f16_col0 are written as IEEE fp16 (FMC_GEMM_F16_TAIL), else INT_MAX
  // LayerNorm folded into the GEMM (fmc_gemm_ln_bf16): A holds the UN-normalised rows, W the weights scaled by gamma;
  // epilogue: acc * rstd[row] - mean[row] * rstd[row] * colsum[n] (+ bias, which carries W beta)
  const float2* rowstats;
  const float* colsum;
  int tiles_m, tiles_n;
  // implicit-GEMM convolution (CONV kernels): A rows are output pixels (n, oh, ow) of a channels-last image, K runs over
  // (ky, kx, cin); the A operand of k-block kb is the input shifted by tap kb / cchunks, fetched row by row with 4-D
  // TMA boxes whose out-of-image part is zero-filled (= the convolution's zero padding)
  int conv_oh, conv_ow, conv_stride, conv_cchunks, conv_rows;  // conv_rows = 128 / OW output rows per tile
};

template <int BN>
struct GemmCfg {
  static constexpr int A_BYTES = GEMM_BM * GEMM_BK * 2;
  static constexpr int B_BYTES = BN * GEMM_BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = (BN <= 128) ? 6 : (BN <= 160 ? 5 : 4);
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024;  // +1024: manual alignment slack
  static constexpr int TMEM_COLS = (2 * BN <= 256) ? 256 : 512;
  static_assert(BN % 16 == 0 && BN >= 16 && BN <= 256, "UMMA N constraint for M=128");
  static_assert(A_BYTES % 1024 == 0 && B_BYTES % 1024 == 0, "SW128 tiles must stay 1024-byte aligned");
};

__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }

template <int BN>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_bf16_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, GemmParams p) {
  pdl_launch_dependents();
  using Cfg = GemmCfg<BN>;
  constexpr int STAGES = Cfg::STAGES;

  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t full_bar[STAGES];
  __shared__ uint64_t empty_bar[STAGES];
  __shared__ uint64_t acc_full_bar[2];
  __shared__ uint64_t acc_empty_bar[2];
  __shared__ uint32_t tmem_base_slot;

  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int num_tiles = p.tiles_m * p.tiles_n;
  const int kblocks = (p.K + GEMM_BK - 1) / GEMM_BK;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&acc_full_bar[a], 1);
      mbar_init(&acc_empty_bar[a], 4);
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc(&tmem_base_slot, Cfg::TMEM_COLS);
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = tmem_base_slot;
  pdl_wait();  // everything above (barriers, TMEM, descriptor prefetch) overlaps the previous kernel's tail

  if (warp == 0) {
    // ------------------------------ TMA producer ------------------------------
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int m_blk = tile / p.tiles_n;
        const int n_blk = tile % p.tiles_n;
        for (int kb = 0; kb < kblocks; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1u);
          mbar_arrive_expect_tx(&full_bar[stage], Cfg::STAGE_BYTES);
          const uint32_t sa = smem_base + stage * Cfg::STAGE_BYTES;
          const uint32_t sb = sa + Cfg::A_BYTES;
          asm volatile(
              "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
              ::"r"(sa), "l"(reinterpret_cast<uint64_t>(&tmA)), "r"(smem_u32(&full_bar[stage])),
              "r"(kb * GEMM_BK), "r"(m_blk * GEMM_BM)
              : "memory");
          asm volatile(
              "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
              ::"r"(sb), "l"(reinterpret_cast<uint64_t>(&tmB)), "r"(smem_u32(&full_bar[stage])),
              "r"(kb * GEMM_BK), "r"(n_blk * BN)
              : "memory");
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1u;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------ MMA issuer ------------------------------
    if (elect_one()) {
      constexpr uint32_t idesc = umma_idesc_bf16(GEMM_BM, BN);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        mbar_wait(&acc_empty_bar[acc], acc_phase ^ 1u);
        tc_fence_after_sync();
        const uint32_t tmem_d = tmem_base + static_cast<uint32_t>(acc * BN);
        for (int kb = 0; kb < kblocks; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after_sync();
          const uint32_t sa = smem_base + stage * Cfg::STAGE_BYTES;
          const uint32_t sb = sa + Cfg::A_BYTES;
          const uint64_t da = umma_desc_k_sw128(sa);
          const uint64_t db = umma_desc_k_sw128(sb);
#pragma unroll
          for (int k = 0; k < GEMM_BK / 16; ++k) {
            // +32 bytes per 16-element K step inside the 128-byte swizzle span (start address is in 16 B units)
            umma_bf16_ss(tmem_d, da + static_cast<uint64_t>(2 * k), db + static_cast<uint64_t>(2 * k), idesc,
                         (kb > 0 || k > 0) ? 1u : 0u);
          }
          umma_commit(&empty_bar[stage]);  // frees the smem stage once these MMAs have read it
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1u;
          }
        }
        umma_commit(&acc_full_bar[acc]);  // accumulator complete -> epilogue
        if (++acc == 2) {
          acc = 0;
          acc_phase ^= 1u;
        }
      }
    }
  } else if (warp >= 4) {
    // ------------------------------ epilogue ------------------------------
    const int q = warp & 3;  // TMEM lane quadrant this warp may access
    const bool geglu = (p.flags & FMC_GEMM_GEGLU) != 0;
    const bool out_f32 = (p.flags & FMC_GEMM_OUT_F32) != 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int m_blk = tile / p.tiles_n;
      const int n_blk = tile % p.tiles_n;
      mbar_wait(&acc_full_bar[acc], acc_phase);
      tc_fence_after_sync();
      const int row = m_blk * GEMM_BM + q * 32 + lane;
      const bool row_ok = row < p.M;
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(acc * BN);
      const float* rb = nullptr;
      if (p.rowbias != nullptr && row_ok) rb = p.rowbias + static_cast<long long>(row / p.rows_per_group) * p.ldrb;

      if (!geglu) {
#pragma unroll 1
        for (int c = 0; c < BN / 16; ++c) {
          const int col0 = n_blk * BN + c * 16;
          if (col0 >= p.N) break;  // warp-uniform
          uint32_t r[16];
          tmem_ld_x16(taddr + static_cast<uint32_t>(c * 16), r);
          tmem_ld_wait();
          float v[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(r[j]);
          if (p.bias != nullptr) {
#pragma unroll
            for (int j = 0; j < 16; j += 4) {
              const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.bias + col0 + j));
              v[j] += b4.x; v[j + 1] += b4.y; v[j + 2] += b4.z; v[j + 3] += b4.w;
            }
          }
          if (row_ok) {
            if (rb != nullptr) {
#pragma unroll
              for (int j = 0; j < 16; j += 4) {
                const float4 b4 = __ldg(reinterpret_cast<const float4*>(rb + col0 + j));
                v[j] += b4.x; v[j + 1] += b4.y; v[j + 2] += b4.z; v[j + 3] += b4.w;
              }
            }
            if (p.residual != nullptr) {
              const uint4* rp = reinterpret_cast<const uint4*>(p.residual + static_cast<long long>(row) * p.ldr + col0);
#pragma unroll
              for (int h = 0; h < 2; ++h) {
                const uint4 u = __ldg(rp + h);
                const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  v[h * 8 + 2 * j] += bf16_lo(w[j]);
                  v[h * 8 + 2 * j + 1] += bf16_hi(w[j]);
                }
              }
            }
            if (out_f32) {
              float4* op = reinterpret_cast<float4*>(static_cast<float*>(p.C) + static_cast<long long>(row) * p.ldc + col0);
#pragma unroll
              for (int j = 0; j < 4; ++j) op[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
            } else {
              uint4* op = reinterpret_cast<uint4*>(static_cast<__nv_bfloat16*>(p.C) + static_cast<long long>(row) * p.ldc + col0);
              op[0] = make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]),
                                 pack_bf16x2(v[6], v[7]));
              op[1] = make_uint4(pack_bf16x2(v[8], v[9]), pack_bf16x2(v[10], v[11]), pack_bf16x2(v[12], v[13]),
                                 pack_bf16x2(v[14], v[15]));
            }
          }
        }
      } else {
        // GEGLU: W rows are interleaved in blocks of 16 (value block, gate block); out = value * gelu(gate).
#pragma unroll 1
        for (int c = 0; c < BN / 32; ++c) {
          const int col0 = n_blk * BN + c * 32;  // column of the value block in the interleaved N space
          if (col0 >= p.N) break;
          uint32_t ra[16], rg[16];
          tmem_ld_x16(taddr + static_cast<uint32_t>(c * 32), ra);
          tmem_ld_x16(taddr + static_cast<uint32_t>(c * 32 + 16), rg);
          tmem_ld_wait();
          float v[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            float a = __uint_as_float(ra[j]);
            float g = __uint_as_float(rg[j]);
            if (p.bias != nullptr) {
              a += __ldg(p.bias + col0 + j);
              g += __ldg(p.bias + col0 + 16 + j);
            }
            v[j] = a * gelu_erf(g);
          }
          const int ocol0 = col0 >> 1;
          if (row_ok) {
            if (p.residual != nullptr) {
              const uint4* rp = reinterpret_cast<const uint4*>(p.residual + static_cast<long long>(row) * p.ldr + ocol0);
#pragma unroll
              for (int h = 0; h < 2; ++h) {
                const uint4 u = __ldg(rp + h);
                const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  v[h * 8 + 2 * j] += bf16_lo(w[j]);
                  v[h * 8 + 2 * j + 1] += bf16_hi(w[j]);
                }
              }
            }
            if (out_f32) {
              float4* op = reinterpret_cast<float4*>(static_cast<float*>(p.C) + static_cast<long long>(row) * p.ldc + ocol0);
#pragma unroll
              for (int j = 0; j < 4; ++j) op[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
            } else {
              uint4* op = reinterpret_cast<uint4*>(static_cast<__nv_bfloat16*>(p.C) + static_cast<long long>(row) * p.ldc + ocol0);
              op[0] = make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]),
                                 pack_bf16x2(v[6], v[7]));
              op[1] = make_uint4(pack_bf16x2(v[8], v[9]), pack_bf16x2(v[10], v[11]), pack_bf16x2(v[12], v[13]),
                                 pack_bf16x2(v[14], v[15]));
            }
          }
        }
      }
      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty_bar[acc]);
      if (++acc == 2) {
        acc = 0;
        acc_phase ^= 1u;
      }
    }
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after_sync();
    tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}


// =====================================================================================================================
// bf16-output GEMM with a shared-memory / TMA epilogue  (the workhorse: every bf16 linear of the step)
//
//   warp 0      TMA producer of the A / W k-slices           (SWIZZLE_128B, BK = 64)
//   warp 1      tcgen05.mma issuer                           (128 x BN x 16 per instruction, fp32 accumulators in TMEM)
//   warp 2      TMEM allocator
//   warp 3      output mover: TMA-loads the residual tile into the staging buffer ahead of time, TMA-stores finished
//               tiles (coalesced 128-byte lines instead of one 32-byte sector per thread)
//   warps 4-11  epilogue: two warps per TMEM lane quadrant, each owning half of the 32-column sub-tiles:
//               tcgen05.ld -> + bias (+ row bias) (+ residual from smem) | value * gelu(gate) -> bf16 -> staging smem
// Accumulators and staging buffers are double-buffered, so MMA of tile i+1, epilogue of tile i and the store of tile
// i-1 overlap.  The staging buffer is a row of [128 x 32] bf16 sub-tiles in SWIZZLE_64B layout (conflict-free for one
// row per thread).
// =====================================================================================================================
constexpr int GEMM2_THREADS = 384;
constexpr int SUB_COLS = 32;                      // output columns per staging sub-tile
constexpr int SUB_BYTES = GEMM_BM * SUB_COLS * 2;  // 8 KB

template <int BN, bool GEGLU, int CL = 1>
struct Gemm2Cfg {
  static constexpr int A_BYTES = GEMM_BM * GEMM_BK * 2;
  static constexpr int B_BYTES = (BN / CL) * GEMM_BK * 2;  // CL = 2: each CTA of the pair stages half of the W slice
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int OUT_COLS = GEGLU ? BN / 2 : BN;
  static constexpr int NSUB = OUT_COLS / SUB_COLS;
  static constexpr int OUT_BYTES = NSUB * SUB_BYTES;
  static constexpr int AVAIL = 227 * 1024 - 1024 - 2 * OUT_BYTES - 512;
  static constexpr int STAGES_RAW = AVAIL / STAGE_BYTES;
  static constexpr int STAGES = STAGES_RAW > 6 ? 6 : STAGES_RAW;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 2 * OUT_BYTES + 1024;
  static constexpr int TMEM_COLS = (2 * BN <= 256) ? 256 : 512;
  static_assert(BN % 16 == 0 && BN >= 32 && BN <= 256, "UMMA N constraint for M=128");
  static_assert(OUT_COLS % SUB_COLS == 0 && NSUB >= 1, "whole sub-tiles");
  static_assert(STAGES >= 3, "not enough shared memory for the k pipeline");
  static_assert(A_BYTES % 1024 == 0 && B_BYTES % 1024 == 0 && OUT_BYTES % 1024 == 0, "1024-byte aligned tiles");
};

// erf-GELU with erf from Abramowitz & Stegun 7.1.28 (|err| <= 3e-7): 1 - erf(x) = (1 + a1 x + ... + a6 x^6)^-16, x >= 0.
// 1 + erf(-x) = 1 - erf(x) is formed directly, so the negative tail has no cancellation.  One MUFU (rcp) per value.
__device__ __forceinline__ float gelu_fast(float g) {
  const float x = fabsf(g) * 0.70710678118654752440f;
  float p = fmaf(x, 0.0000430638f, 0.0002765672f);
  p = fmaf(x, p, 0.0001520143f);
  p = fmaf(x, p, 0.0092705272f);
  p = fmaf(x, p, 0.0422820123f);
  p = fmaf(x, p, 0.0705230784f);
  p = fmaf(x, p, 1.0f);
  float r = rcp_fast(p);
  r *= r; r *= r; r *= r; r *= r;            // (1/p)^16 = 1 - erf(x)
  const float one_plus_erf = g >= 0.f ? 2.0f - r : r;
  return 0.5f * g * one_plus_erf;
}

// Two GEGLU outputs at once, a * gelu_erf(g), on packed fp32 pairs (same polynomial as gelu_fast): the level-0 GEGLU
// GEMM (K = 320) is bound by its epilogue's issue slots, and the pair form needs ~12 instructions per output instead
// of ~20.  1 + erf(g / sqrt 2) = 1 + copysign(1 - r, g) with r = (1 / p(|g| / sqrt 2))^16.
__device__ __forceinline__ void geglu_pair(float a0, float a1, float g0, float g1, float& o0, float& o1) {
  const uint64_t g2 = f2_pack(g0, g1);
  const uint64_t x2 = f2_mul(g2, f2_pack(0.70710678118654752440f, 0.70710678118654752440f)) & 0x7FFFFFFF7FFFFFFFull;
  uint64_t p2 = f2_fma(x2, f2_pack(0.0000430638f, 0.0000430638f), f2_pack(0.0002765672f, 0.0002765672f));
  p2 = f2_fma(x2, p2, f2_pack(0.0001520143f, 0.0001520143f));
  p2 = f2_fma(x2, p2, f2_pack(0.0092705272f, 0.0092705272f));
  p2 = f2_fma(x2, p2, f2_pack(0.0422820123f, 0.0422820123f));
  p2 = f2_fma(x2, p2, f2_pack(0.0705230784f, 0.0705230784f));
  p2 = f2_fma(x2, p2, f2_pack(1.0f, 1.0f));
  float p0, p1;
  f2_unpack(p2, p0, p1);
  uint64_t r2 = f2_pack(rcp_fast(p0), rcp_fast(p1));
  r2 = f2_mul(r2, r2); r2 = f2_mul(r2, r2); r2 = f2_mul(r2, r2); r2 = f2_mul(r2, r2);   // 1 - erf(|x|)
  const uint64_t s2 = f2_fma(r2, f2_pack(-1.0f, -1.0f), f2_pack(1.0f, 1.0f));            // erf(|x|) >= 0
  const uint64_t e2 = s2 | (g2 & 0x8000000080000000ull);                                 // erf(x)
  const uint64_t h2 = f2_fma(e2, f2_pack(0.5f, 0.5f), f2_pack(0.5f, 0.5f));              // (1 + erf) / 2
  const uint64_t o2 = f2_mul(f2_mul(f2_pack(a0, a1), g2), h2);
  f2_unpack(o2, o0, o1);
}

__device__ __forceinline__ void ld_shared_v4(uint32_t addr, uint32_t (&v)[4]) {
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]) : "r"(addr));
}
__device__ __forceinline__ void epi_bar_arrive_warp(uint64_t* bar, int lane) {
  __syncwarp();
  if (lane == 0) mbar_arrive(bar);
}

// CL = 2: CTA-pair MMA (tcgen05 cta_group::2).  The two CTAs of a cluster own vertically adjacent output tiles
// (m_blk = 2 i + rank, same n_blk) and execute ONE 256 x BN MMA stream issued by the even CTA: every CTA stages its own
// 128 A rows and only HALF of the W slice (BN / 2 rows), the tensor cores of both SMs read the two halves.  Per SM and
// k-block that is 128 + BN / 2 staged rows instead of 128 + BN: with both operands in shared memory a 128 x 256 tile
// needs ~190 B/clk of smem traffic at the full MMA rate against ~128 B/clk available, which is what caps the 1-CTA
// kernel near 57 % of the tensor peak on the K >= 1280 shapes.
//   full barrier   : the even CTA's; it expects the bytes of all four loads (both CTAs' TMA signal it)
//   empty barrier  : one per CTA, released by the issuing thread's commit, multicast to both CTAs
//   acc_full       : commit multicast to both CTAs (each epilogue drains its own 128 rows from its own TMEM)
//   acc_empty      : the even CTA's, 16 arrivals (the odd CTA's epilogue warps arrive remotely)
template <int BN, bool GEGLU, int CL, bool CONV = false>
__global__ void __launch_bounds__(GEMM2_THREADS, 1)
gemm_bf16_tma_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                     const __grid_constant__ CUtensorMap tmC, const __grid_constant__ CUtensorMap tmR, GemmParams p) {
  pdl_launch_dependents();
  using Cfg = Gemm2Cfg<BN, GEGLU, CL>;
  static_assert(CL == 1 || CL == 2, "cluster size");
  static_assert(((BN / 2) * 128) % 1024 == 0, "half W slices must stay 1024-byte aligned");
  constexpr int STAGES = Cfg::STAGES;
  constexpr int NSUB = Cfg::NSUB;

  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t full_bar[STAGES];
  __shared__ uint64_t empty_bar[STAGES];
  __shared__ uint64_t acc_full_bar[2];
  __shared__ uint64_t acc_empty_bar[2];
  __shared__ uint64_t buf_ready_bar[2];  // staging buffer b holds the residual (or is simply free) for its next tile
  __shared__ uint64_t out_ready_bar[2];  // staging buffer b holds a finished output tile
  __shared__ uint32_t tmem_base_slot;

  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t out_base = smem_base + STAGES * Cfg::STAGE_BYTES;
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int kblocks = (p.K + GEMM_BK - 1) / GEMM_BK;
  const bool has_res = p.residual != nullptr;
  const int n_out = GEGLU ? p.N / 2 : p.N;
  // work units: CL == 1: output tiles; CL == 2: pairs of vertically adjacent tiles, one per CTA of the cluster
  const int rank = CL == 2 ? static_cast<int>(cluster_ctarank()) : 0;
  const int unit0 = static_cast<int>(blockIdx.x) / CL;
  const int unit_step = static_cast<int>(gridDim.x) / CL;
  const int num_tiles = ((p.tiles_m + CL - 1) / CL) * p.tiles_n;
  // unit -> (m_blk, n_blk) of THIS CTA; an odd last row of tiles is computed by both CTAs (identical stores)
  auto m_blk_of = [&](int unit) -> int { const int m = (unit / p.tiles_n) * CL + rank; return m < p.tiles_m ? m : p.tiles_m - 1; };

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    tma_prefetch_desc(&tmC);
    if (has_res) tma_prefetch_desc(&tmR);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&acc_full_bar[a], 1);
      mbar_init(&acc_empty_bar[a], 8 * CL);
      mbar_init(&buf_ready_bar[a], 1);
      mbar_init(&out_ready_bar[a], 8);
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    if constexpr (CL == 2) tmem_alloc_2sm(&tmem_base_slot, Cfg::TMEM_COLS);
    else tmem_alloc(&tmem_base_slot, Cfg::TMEM_COLS);
  }
  tc_fence_before_sync();
  __syncthreads();
  if constexpr (CL == 2) cluster_sync_all();  // the peer's barriers / TMEM exist before anything is sent to them
  tc_fence_after_sync();
  const uint32_t tmem_base = tmem_base_slot;
  pdl_wait();  // everything above (barriers, TMEM, descriptor prefetch) overlaps the previous kernel's tail

  if (warp == 0) {
    // ------------------------------ TMA producer (A, W) ------------------------------
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = unit0; tile < num_tiles; tile += unit_step) {
        const int m_blk = m_blk_of(tile);
        const int n_blk = tile % p.tiles_n;
        for (int kb = 0; kb < kblocks; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1u);
          const uint32_t sa = smem_base + stage * Cfg::STAGE_BYTES;
          // CONV: tap (dy, dx) and 64-channel chunk of this k-block
          const int tap = CONV ? kb / p.conv_cchunks : 0;
          const int c0 = CONV ? (kb - tap * p.conv_cchunks) * GEMM_BK : 0;
          const int dy = tap / 3, dx = tap - dy * 3;
          if constexpr (CL == 2) {
            // own A rows + own half of the W slice (tmB's box is BN / 2 rows) into own smem; the bytes of both CTAs
            // are counted by the even CTA's full barrier
            if (rank == 0) mbar_arrive_expect_tx(&full_bar[stage], 2 * Cfg::STAGE_BYTES);
            const uint32_t fb = mapa_shared(smem_u32(&full_bar[stage]), 0);
            if constexpr (CONV) {
              for (int j = 0; j < p.conv_rows; ++j) {
                const int pr = m_blk * p.conv_rows + j;  // output row (n, oh) of the flattened image stack
                const int n = pr / p.conv_oh, oh = pr - n * p.conv_oh;
                tma_load_4d_2sm(sa + j * p.conv_ow * 128, &tmA, fb, c0, dx - 1, oh * p.conv_stride + dy - 1, n);
              }
            } else {
              tma_load_2d_2sm(sa, &tmA, fb, kb * GEMM_BK, m_blk * GEMM_BM);
            }
            tma_load_2d_2sm(sa + Cfg::A_BYTES, &tmB, fb, kb * GEMM_BK, n_blk * BN + rank * (BN / 2));
          } else {
            mbar_arrive_expect_tx(&full_bar[stage], Cfg::STAGE_BYTES);
            if constexpr (CONV) {
              for (int j = 0; j < p.conv_rows; ++j) {
                const int pr = m_blk * p.conv_rows + j;
                const int n = pr / p.conv_oh, oh = pr - n * p.conv_oh;
                tma_load_4d_a(sa + j * p.conv_ow * 128, &tmA, &full_bar[stage], c0, dx - 1, oh * p.conv_stride + dy - 1, n);
              }
            } else {
              tma_load_2d_a(sa, &tmA, &full_bar[stage], kb * GEMM_BK, m_blk * GEMM_BM);
            }
            tma_load_2d_a(sa + Cfg::A_BYTES, &tmB, &full_bar[stage], kb * GEMM_BK, n_blk * BN);
          }
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1u;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------ MMA issuer ------------------------------
    // CL = 2: only the even CTA of the pair issues (M = 256); its commits reach both CTAs.
    if ((CL == 1 || rank == 0) && elect_one()) {
      constexpr uint32_t idesc = umma_idesc_bf16(GEMM_BM * CL, BN);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = unit0; tile < num_tiles; tile += unit_step) {
        mbar_wait(&acc_empty_bar[acc], acc_phase ^ 1u);
        tc_fence_after_sync();
        const uint32_t tmem_d = tmem_base + static_cast<uint32_t>(acc * BN);
        for (int kb = 0; kb < kblocks; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after_sync();
          const uint32_t sa = smem_base + stage * Cfg::STAGE_BYTES;
          const uint64_t da = umma_desc_k_sw128(sa);
          const uint64_t db = umma_desc_k_sw128(sa + Cfg::A_BYTES);
#pragma unroll
          for (int k = 0; k < GEMM_BK / 16; ++k) {
            if constexpr (CL == 2)
              umma_bf16_ss_2sm(tmem_d, da + static_cast<uint64_t>(2 * k), db + static_cast<uint64_t>(2 * k), idesc,
                               (kb > 0 || k > 0) ? 1u : 0u);
            else
              umma_bf16_ss(tmem_d, da + static_cast<uint64_t>(2 * k), db + static_cast<uint64_t>(2 * k), idesc,
                           (kb > 0 || k > 0) ? 1u : 0u);
          }
          if constexpr (CL == 2) umma_commit_2sm(&empty_bar[stage], static_cast<uint16_t>(3));
          else umma_commit(&empty_bar[stage]);
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1u;
          }
        }
        if constexpr (CL == 2) umma_commit_2sm(&acc_full_bar[acc], static_cast<uint16_t>(3));
        else umma_commit(&acc_full_bar[acc]);
        if (++acc == 2) {
          acc = 0;
          acc_phase ^= 1u;
        }
      }
    }
  } else if (warp == 3) {
    // ------------------------------ output mover ------------------------------
    if (elect_one()) {
      // arm staging buffer `b` for tile `tile`: residual tile lands there (complete_tx) or it is simply marked free
      auto arm = [&](int b, int tile) {
        if (has_res) {
          const int m_blk = m_blk_of(tile);
          const int n_blk = tile % p.tiles_n;
          int nsub_ok = 0;
          for (int s = 0; s < NSUB; ++s) nsub_ok += (n_blk * Cfg::OUT_COLS + s * SUB_COLS < n_out) ? 1 : 0;
          mbar_arrive_expect_tx(&buf_ready_bar[b], static_cast<uint32_t>(nsub_ok * SUB_BYTES));
          for (int s = 0; s < nsub_ok; ++s) {
            tma_load_2d_a(out_base + b * Cfg::OUT_BYTES + s * SUB_BYTES, &tmR, &buf_ready_bar[b],
                          n_blk * Cfg::OUT_COLS + s * SUB_COLS, m_blk * GEMM_BM);
          }
        } else {
          mbar_arrive(&buf_ready_bar[b]);
        }
      };
      int tile = unit0;
      if (tile < num_tiles) arm(0, tile);
      if (tile + unit_step < num_tiles) arm(1, tile + unit_step);
      int it = 0;
      for (; tile < num_tiles; tile += unit_step, ++it) {
        const int b = it & 1;
        const uint32_t ph = static_cast<uint32_t>(it >> 1) & 1u;
        const int m_blk = m_blk_of(tile);
        const int n_blk = tile % p.tiles_n;
        mbar_wait(&out_ready_bar[b], ph);
        for (int s = 0; s < NSUB; ++s) {
          const int col = n_blk * Cfg::OUT_COLS + s * SUB_COLS;
          if (col < n_out) {
            asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                             reinterpret_cast<uint64_t>(&tmC)),
                         "r"(out_base + b * Cfg::OUT_BYTES + s * SUB_BYTES), "r"(col), "r"(m_blk * GEMM_BM)
                         : "memory");
          }
        }
        tma_store_commit();
        const int next2 = tile + 2 * unit_step;
        if (next2 < num_tiles) {
          tma_store_wait_read<0>();  // the store has drained buffer b: reuse it for the tile after next
          arm(b, next2);
        }
      }
      tma_store_wait<0>();
    }
  } else if (warp >= 4) {
    // ------------------------------ epilogue ------------------------------
    const int q = warp & 3;                   // TMEM lane quadrant of this warp
    const int half = (warp - 4) >> 2;         // which half of the sub-tiles
    constexpr int S_SPLIT = (NSUB + 1) / 2;
    const int s_begin = half == 0 ? 0 : S_SPLIT;
    const int s_end = half == 0 ? S_SPLIT : NSUB;
    const int r_in_tile = q * 32 + lane;
    const uint32_t row_off = static_cast<uint32_t>(r_in_tile) * 64u;
    const uint32_t sw = static_cast<uint32_t>((r_in_tile >> 1) & 3);
    int it = 0;
    for (int tile = unit0; tile < num_tiles; tile += unit_step, ++it) {
      const int b = it & 1;
      const uint32_t ph = static_cast<uint32_t>(it >> 1) & 1u;
      const int m_blk = m_blk_of(tile);
      const int n_blk = tile % p.tiles_n;
      const int row = m_blk * GEMM_BM + r_in_tile;
      const float* rb = nullptr;
      if (p.rowbias != nullptr && row < p.M) rb = p.rowbias + static_cast<long long>(row / p.rows_per_group) * p.ldrb;
      float ln_rstd = 1.f, ln_shift = 0.f;  // folded LayerNorm of this row: acc * rstd - mean * rstd * colsum[n]
      if (p.rowstats != nullptr && row < p.M) {
        const float2 st = __ldg(p.rowstats + row);
        ln_rstd = st.y;
        ln_shift = -st.x * st.y;
      }
      mbar_wait(&acc_full_bar[b], ph);
      mbar_wait(&buf_ready_bar[b], ph);
      tc_fence_after_sync();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(b * BN);
      const uint32_t buf = out_base + b * Cfg::OUT_BYTES + row_off;
#pragma unroll 1
      for (int s = s_begin; s < s_end; ++s) {
        const int ocol0 = n_blk * Cfg::OUT_COLS + s * SUB_COLS;  // first output column of this sub-tile
        if (ocol0 >= n_out) break;                                // warp-uniform
        float v[32];
        if constexpr (!GEGLU) {
          uint32_t r[32];
          tmem_ld_x32(taddr + static_cast<uint32_t>(s * 32), r);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
          if (p.rowstats != nullptr) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              const float4 s4 = __ldg(reinterpret_cast<const float4*>(p.colsum + ocol0 + j));
              v[j] = fmaf(v[j], ln_rstd, ln_shift * s4.x); v[j + 1] = fmaf(v[j + 1], ln_rstd, ln_shift * s4.y);
              v[j + 2] = fmaf(v[j + 2], ln_rstd, ln_shift * s4.z); v[j + 3] = fmaf(v[j + 3], ln_rstd, ln_shift * s4.w);
            }
          }
          if (p.bias != nullptr) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.bias + ocol0 + j));
              v[j] += b4.x; v[j + 1] += b4.y; v[j + 2] += b4.z; v[j + 3] += b4.w;
            }
          }
          if (rb != nullptr) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              const float4 b4 = __ldg(reinterpret_cast<const float4*>(rb + ocol0 + j));
              v[j] += b4.x; v[j + 1] += b4.y; v[j + 2] += b4.z; v[j + 3] += b4.w;
            }
          }
        } else {
          // interleaved N space: [16 value | 16 gate] blocks; 64 accumulator columns -> 32 outputs
          uint32_t r0[32], r1[32];
          tmem_ld_x32(taddr + static_cast<uint32_t>(s * 64), r0);
          tmem_ld_x32(taddr + static_cast<uint32_t>(s * 64 + 32), r1);
          tmem_ld_wait();
          const int ncol0 = n_blk * BN + s * 64;
#pragma unroll
          for (int hblk = 0; hblk < 2; ++hblk) {
            const uint32_t(&rr)[32] = hblk == 0 ? r0 : r1;
            float bv[32];
            if (p.bias != nullptr) {
#pragma unroll
              for (int j = 0; j < 32; j += 4) {
                const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.bias + ncol0 + hblk * 32 + j));
                bv[j] = b4.x; bv[j + 1] = b4.y; bv[j + 2] = b4.z; bv[j + 3] = b4.w;
              }
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j) bv[j] = 0.f;
            }
            float acc[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) acc[j] = __uint_as_float(rr[j]);
            if (p.rowstats != nullptr) {
#pragma unroll
              for (int j = 0; j < 32; j += 4) {
                const float4 s4 = __ldg(reinterpret_cast<const float4*>(p.colsum + ncol0 + hblk * 32 + j));
                acc[j] = fmaf(acc[j], ln_rstd, ln_shift * s4.x); acc[j + 1] = fmaf(acc[j + 1], ln_rstd, ln_shift * s4.y);
                acc[j + 2] = fmaf(acc[j + 2], ln_rstd, ln_shift * s4.z); acc[j + 3] = fmaf(acc[j + 3], ln_rstd, ln_shift * s4.w);
              }
            }
#pragma unroll
            for (int j = 0; j < 16; j += 2)
              geglu_pair(acc[j] + bv[j], acc[j + 1] + bv[j + 1], acc[16 + j] + bv[16 + j], acc[17 + j] + bv[17 + j],
                         v[hblk * 16 + j], v[hblk * 16 + j + 1]);
          }
        }
        const uint32_t sub = buf + static_cast<uint32_t>(s * SUB_BYTES);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const uint32_t addr = sub + ((static_cast<uint32_t>(c) ^ sw) << 4);
          if (has_res) {
            uint32_t w[4];
            ld_shared_v4(addr, w);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              v[c * 8 + 2 * j] += bf16_lo(w[j]);
              v[c * 8 + 2 * j + 1] += bf16_hi(w[j]);
            }
          }
          if (ocol0 >= p.f16_col0)  // warp-uniform: this 32-column sub-tile belongs to the fp16 tail
            st_shared_v4(addr, pack_f16x2(v[c * 8], v[c * 8 + 1]), pack_f16x2(v[c * 8 + 2], v[c * 8 + 3]),
                         pack_f16x2(v[c * 8 + 4], v[c * 8 + 5]), pack_f16x2(v[c * 8 + 6], v[c * 8 + 7]));
          else
            st_shared_v4(addr, pack_bf16x2(v[c * 8], v[c * 8 + 1]), pack_bf16x2(v[c * 8 + 2], v[c * 8 + 3]),
                         pack_bf16x2(v[c * 8 + 4], v[c * 8 + 5]), pack_bf16x2(v[c * 8 + 6], v[c * 8 + 7]));
        }
      }
      tc_fence_before_sync();
      fence_proxy_async_smem();  // staging writes -> visible to the TMA store
      if constexpr (CL == 2) {
        // the accumulator pair is released to the issuing (even) CTA
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(mapa_shared(smem_u32(&acc_empty_bar[b]), 0));
      } else {
        epi_bar_arrive_warp(&acc_empty_bar[b], lane);
      }
      if (lane == 0) mbar_arrive(&out_ready_bar[b]);
    }
  }

  tc_fence_before_sync();
  __syncthreads();
  if constexpr (CL == 2) cluster_sync_all();  // no CTA leaves (or frees TMEM) while the pair's MMAs / arrivals are in flight
  if (warp == 2) {
    tc_fence_after_sync();
    if constexpr (CL == 2) tmem_dealloc_2sm(tmem_base, Cfg::TMEM_COLS);
    else tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

template <int BN, bool GEGLU, int CL, bool CONV = false>
static int launch_gemm2(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmC, const CUtensorMap& tmR,
                        GemmParams& p, cudaStream_t stream) {
  using Cfg = Gemm2Cfg<BN, GEGLU, CL>;
  static unsigned long long attr_devs = 0;  // per device: the attribute belongs to the (device, function) pair
  if (first_use_on_this_device(&attr_devs)) {
    FMC_CUDA_OK(cudaFuncSetAttribute(gemm_bf16_tma_kernel<BN, GEGLU, CL, CONV>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     Cfg::SMEM_BYTES));
  }
  p.tiles_m = ceil_div(p.M, GEMM_BM);
  p.tiles_n = ceil_div(p.N, BN);
  const int units = ceil_div(p.tiles_m, CL) * p.tiles_n;
  const int max_units = device_sm_count() / CL;
  const int grid = (units < max_units ? units : max_units) * CL;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(static_cast<unsigned>(grid));
  cfg.blockDim = dim3(GEMM2_THREADS);
  cfg.dynamicSmemBytes = Cfg::SMEM_BYTES;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  int na = 0;
  if (CL > 1) {
    attr[na].id = cudaLaunchAttributeClusterDimension;
    attr[na].val.clusterDim.x = CL;
    attr[na].val.clusterDim.y = 1;
    attr[na].val.clusterDim.z = 1;
    ++na;
  }
  if (pdl_enabled()) {
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  cfg.attrs = attr;
  cfg.numAttrs = na;
  FMC_CUDA_OK(cudaLaunchKernelEx(&cfg, gemm_bf16_tma_kernel<BN, GEGLU, CL, CONV>, tmA, tmB, tmC, tmR, p));
  return check_launch("gemm_bf16_tma_kernel");
}

template <int BN>
static int launch_gemm(const CUtensorMap& tmA, const CUtensorMap& tmB, GemmParams& p, cudaStream_t stream) {
  using Cfg = GemmCfg<BN>;
  static unsigned long long attr_devs = 0;  // per device: the attribute belongs to the (device, function) pair
  if (first_use_on_this_device(&attr_devs)) {
    FMC_CUDA_OK(cudaFuncSetAttribute(gemm_bf16_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
  }
  p.tiles_m = ceil_div(p.M, GEMM_BM);
  p.tiles_n = ceil_div(p.N, BN);
  const int tiles = p.tiles_m * p.tiles_n;
  const int grid = tiles < device_sm_count() ? tiles : device_sm_count();
  launch_k(gemm_bf16_kernel<BN>, dim3(grid), dim3(GEMM_THREADS), Cfg::SMEM_BYTES, stream, tmA, tmB, p);
  return check_launch("gemm_bf16_kernel");
}

}  // namespace fmc

using namespace fmc;

static int gemm_impl(const void* A, long long lda, const void* W, long long ldw, void* C, long long ldc, int M, int N, int K,
                     const float* bias, const void* residual, long long ldr, const float* rowbias, int rows_per_group,
                     long long ldrb, int flags, int tile_n, const float2* rowstats, const float* colsum, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  FMC_REQUIRE(A && W && C, FMC_ERR_ARG, "fmc_gemm_bf16: null operand");
  FMC_REQUIRE(M > 0 && N > 0 && K > 0, FMC_ERR_SHAPE, "fmc_gemm_bf16: empty problem %dx%dx%d", M, N, K);
  FMC_REQUIRE(K % 8 == 0 && lda % 8 == 0 && ldw % 8 == 0, FMC_ERR_SHAPE,
              "fmc_gemm_bf16: K=%d, lda=%lld, ldw=%lld must be multiples of 8 (16-byte TMA rows)", K, lda, ldw);
  const bool geglu = (flags & FMC_GEMM_GEGLU) != 0;
  FMC_REQUIRE(N % (geglu ? 32 : 16) == 0, FMC_ERR_SHAPE, "fmc_gemm_bf16: N=%d must be a multiple of %d", N,
              geglu ? 32 : 16);
  const int align_out = (flags & FMC_GEMM_OUT_F32) ? 4 : 8;
  FMC_REQUIRE(ldc % align_out == 0 && (reinterpret_cast<uintptr_t>(C) & 15) == 0, FMC_ERR_SHAPE,
              "fmc_gemm_bf16: output must be 16-byte aligned per row (ldc=%lld)", ldc);
  FMC_REQUIRE(residual == nullptr || (ldr % 8 == 0 && (reinterpret_cast<uintptr_t>(residual) & 15) == 0),
              FMC_ERR_SHAPE, "fmc_gemm_bf16: residual must be 16-byte aligned per row (ldr=%lld)", ldr);
  FMC_REQUIRE(rowbias == nullptr || (rows_per_group > 0 && ldrb % 4 == 0), FMC_ERR_ARG,
              "fmc_gemm_bf16: rowbias needs rows_per_group > 0 and ldrb %% 4 == 0");

  const bool out_f32 = (flags & FMC_GEMM_OUT_F32) != 0;
  const int n_out = geglu ? N / 2 : N;
  const bool tma_epilogue = !out_f32 && n_out % SUB_COLS == 0 && (!geglu || N % 64 == 0);

  int bn = tile_n;
  if (tma_epilogue) {
    if (geglu) {
      if (bn != 128 && bn != 256) bn = (N % 256 == 0 || N > 1024) ? 256 : 128;
    } else if (bn != 64 && bn != 128 && bn != 160 && bn != 256) {
      bn = N <= 64 ? 64 : (N <= 128 ? 128 : 160);
    }
  } else {
    if (bn == 0) bn = (N % 160 == 0) ? 160 : ((N % 128 == 0 || N > 128) ? 128 : 64);
    FMC_REQUIRE(bn == 64 || bn == 128 || bn == 160 || bn == 256, FMC_ERR_SHAPE, "fmc_gemm_bf16: unsupported tile_n %d", bn);
  }

  // CTA-pair MMA (cta_group::2, see gemm_bf16_tma_kernel); FMC_GEMM_1CTA=1 forces the single-CTA kernel (A/B runs).
  // (An earlier variant that only multicast the W slices to two independent CTAs measured within +-3 % of the plain
  // kernel on all twelve level 0-2 shapes: profiles/r01_gemm_cluster_multicast.txt.)
  // Measured per shape (profiles/r01_gemm_2cta.txt): the pair wins 5-15 % from K = 640 up, and loses 7-23 % at K = 320,
  // where a tile is only 20 MMAs and the cross-CTA hand-offs per tile dominate.
  static const bool cluster_allowed = getenv("FMC_GEMM_1CTA") == nullptr;
  static const bool cluster_forced = getenv("FMC_GEMM_2CTA") != nullptr;
  const bool use_cluster = tma_epilogue && cluster_allowed && M > GEMM_BM &&
                           (cluster_forced || K >= 1280 || (K >= 640 && N > 640));
  if (tma_epilogue && !geglu && tile_n == 0 && N > 128) {
    // wave quantisation: with few row tiles (levels 2-3) 128-wide tiles can need fewer / cheaper rounds than 160-wide
    const int cl = use_cluster ? 2 : 1;
    const int slots = device_sm_count() / cl;
    const int rows = ceil_div(ceil_div(M, GEMM_BM), cl);
    const long long cost160 = static_cast<long long>(ceil_div(rows * ceil_div(N, 160), slots)) * 160;
    const long long cost128 = static_cast<long long>(ceil_div(rows * ceil_div(N, 128), slots)) * 128;
    bn = (cost128 * 10 < cost160 * 9) ? 128 : 160;
  }
  if (tma_epilogue && !geglu && bn == 256 && !use_cluster) bn = 160;  // 128 x 256 bf16 staging only fits the pair kernel

  CUtensorMap tmA, tmB;
  {
    const uint64_t dims[2] = {static_cast<uint64_t>(K), static_cast<uint64_t>(M)};
    const uint64_t strides[1] = {static_cast<uint64_t>(lda) * 2};
    const uint32_t box[2] = {GEMM_BK, GEMM_BM};
    int rc = make_tmap_bf16(&tmA, A, 2, dims, strides, box, true);
    if (rc != FMC_OK) return rc;
  }
  {
    const uint64_t dims[2] = {static_cast<uint64_t>(K), static_cast<uint64_t>(N)};
    const uint64_t strides[1] = {static_cast<uint64_t>(ldw) * 2};
    const uint32_t box[2] = {GEMM_BK, static_cast<uint32_t>(use_cluster ? bn / 2 : bn)};
    int rc = make_tmap_bf16(&tmB, W, 2, dims, strides, box, true);
    if (rc != FMC_OK) return rc;
  }
  GemmParams p{};
  p.M = M; p.N = N; p.K = K;
  p.C = C; p.ldc = ldc;
  p.bias = bias;
  p.residual = static_cast<const __nv_bfloat16*>(residual); p.ldr = ldr;
  p.rowbias = rowbias; p.rows_per_group = rows_per_group > 0 ? rows_per_group : 1; p.ldrb = ldrb;
  p.flags = flags;
  p.rowstats = rowstats;
  p.colsum = colsum;
  FMC_REQUIRE(rowstats == nullptr || (tma_epilogue && colsum != nullptr), FMC_ERR_ARG,
              "fmc_gemm_ln_bf16: the folded LayerNorm needs the bf16 TMA-epilogue path and a column-sum vector");
  p.f16_col0 = 0x7FFFFFFF;
  if ((flags & FMC_GEMM_F16_TAIL) != 0) {
    const int col0 = (flags >> 8) * 32;
    FMC_REQUIRE(tma_epilogue && !geglu && residual == nullptr && col0 < N, FMC_ERR_ARG,
                "fmc_gemm_bf16: FMC_GEMM_F16_TAIL needs the bf16 TMA-epilogue path without GEGLU / residual");
    p.f16_col0 = col0;
  }
  if (tma_epilogue) {
    CUtensorMap tmC, tmR;
    const uint32_t box[2] = {SUB_COLS, GEMM_BM};
    {
      const uint64_t dims[2] = {static_cast<uint64_t>(n_out), static_cast<uint64_t>(M)};
      const uint64_t strides[1] = {static_cast<uint64_t>(ldc) * 2};
      int rc = make_tmap_bf16_sw(&tmC, C, 2, dims, strides, box, 64);
      if (rc != FMC_OK) return rc;
    }
    if (residual != nullptr) {
      const uint64_t dims[2] = {static_cast<uint64_t>(n_out), static_cast<uint64_t>(M)};
      const uint64_t strides[1] = {static_cast<uint64_t>(ldr) * 2};
      int rc = make_tmap_bf16_sw(&tmR, residual, 2, dims, strides, box, 64);
      if (rc != FMC_OK) return rc;
    } else {
      tmR = tmC;
    }
    if (use_cluster) {
      if (geglu) {
        if (bn == 256) return launch_gemm2<256, true, 2>(tmA, tmB, tmC, tmR, p, stream);
        return launch_gemm2<128, true, 2>(tmA, tmB, tmC, tmR, p, stream);
      }
      switch (bn) {
        case 64: return launch_gemm2<64, false, 2>(tmA, tmB, tmC, tmR, p, stream);
        case 128: return launch_gemm2<128, false, 2>(tmA, tmB, tmC, tmR, p, stream);
        case 256: return launch_gemm2<256, false, 2>(tmA, tmB, tmC, tmR, p, stream);
        default: return launch_gemm2<160, false, 2>(tmA, tmB, tmC, tmR, p, stream);
      }
    }
    if (geglu) {
      if (bn == 256) return launch_gemm2<256, true, 1>(tmA, tmB, tmC, tmR, p, stream);
      return launch_gemm2<128, true, 1>(tmA, tmB, tmC, tmR, p, stream);
    }
    switch (bn) {
      case 64: return launch_gemm2<64, false, 1>(tmA, tmB, tmC, tmR, p, stream);
      case 128: return launch_gemm2<128, false, 1>(tmA, tmB, tmC, tmR, p, stream);
      default: return launch_gemm2<160, false, 1>(tmA, tmB, tmC, tmR, p, stream);
    }
  }
  switch (bn) {
    case 64: return launch_gemm<64>(tmA, tmB, p, stream);
    case 128: return launch_gemm<128>(tmA, tmB, p, stream);
    case 160: return launch_gemm<160>(tmA, tmB, p, stream);
    default: return launch_gemm<256>(tmA, tmB, p, stream);
  }
}


extern "C" int fmc_gemm_bf16(const void* A, long long lda, const void* W, long long ldw, void* C, long long ldc, int M,
                             int N, int K, const float* bias, const void* residual, long long ldr,
                             const float* rowbias, int rows_per_group, long long ldrb, int flags, int tile_n,
                             void* stream_) {
  return gemm_impl(A, lda, W, ldw, C, ldc, M, N, K, bias, residual, ldr, rowbias, rows_per_group, ldrb, flags, tile_n,
                   nullptr, nullptr, stream_);
}

// C = LayerNorm(A) Wo^T + b with the normalisation folded into the GEMM: W = Wo * gamma (per input channel), colsum[n] =
// sum_k W[n, k], bias = b + Wo beta, rowstats[row] = (mean, rstd) of A's rows (fmc_rowstats_bf16).
extern "C" int fmc_gemm_ln_bf16(const void* A, long long lda, const void* W, long long ldw, void* C, long long ldc, int M,
                                int N, int K, const float* bias, const float* colsum, const void* rowstats, int flags,
                                int tile_n, void* stream_) {
  FMC_REQUIRE(colsum != nullptr && rowstats != nullptr, FMC_ERR_ARG, "fmc_gemm_ln_bf16: null colsum / rowstats");
  return gemm_impl(A, lda, W, ldw, C, ldc, M, N, K, bias, nullptr, 0, nullptr, 0, 0, flags, tile_n,
                   static_cast<const float2*>(rowstats), colsum, stream_);
}

// 3x3 convolution (padding 1, stride 1 or 2) on channels-last bf16 images as an implicit GEMM on the same tcgen05 kernel:
// out[n, oh, ow, :] = sum_{ky,kx,c} x[n, oh*s + ky - 1, ow*s + kx - 1, c] * w[:, ky, kx, c] (+ bias) (+ residual).
extern "C" int fmc_conv3x3_bf16(const void* X, const void* W, void* Out, const float* bias, const void* residual,
                                int images, int H, int Wd, int Cin, int Cout, int stride, int tile_n, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  FMC_REQUIRE(X && W && Out, FMC_ERR_ARG, "fmc_conv3x3_bf16: null operand");
  FMC_REQUIRE(stride == 1 || stride == 2, FMC_ERR_SHAPE, "fmc_conv3x3_bf16: stride %d not in {1, 2}", stride);
  FMC_REQUIRE(images > 0 && H > 0 && Wd > 0 && H % stride == 0 && Wd % stride == 0, FMC_ERR_SHAPE,
              "fmc_conv3x3_bf16: image %d x %d x %d not divisible by the stride", images, H, Wd);
  const int OH = H / stride, OW = Wd / stride;
  FMC_REQUIRE(Cin % 64 == 0 && Cout % 32 == 0, FMC_ERR_SHAPE,
              "fmc_conv3x3_bf16: Cin=%d must be a multiple of 64 and Cout=%d of 32", Cin, Cout);
  FMC_REQUIRE(OW <= 128 && 128 % OW == 0 && OW >= 4, FMC_ERR_SHAPE,
              "fmc_conv3x3_bf16: output width %d must divide 128 (a 128-pixel tile is 128 / OW whole output rows)", OW);
  FMC_REQUIRE((reinterpret_cast<uintptr_t>(Out) & 15) == 0 && (reinterpret_cast<uintptr_t>(X) & 15) == 0, FMC_ERR_SHAPE,
              "fmc_conv3x3_bf16: tensors must be 16-byte aligned");
  const int M = images * OH * OW, N = Cout, K = 9 * Cin;

  static const bool cluster_allowed = getenv("FMC_GEMM_1CTA") == nullptr;
  const bool use_cluster = cluster_allowed && M > GEMM_BM;
  int bn = tile_n;
  if (bn != 128 && bn != 160) {
    const int cl = use_cluster ? 2 : 1;
    const int slots = device_sm_count() / cl;
    const int rows = ceil_div(ceil_div(M, GEMM_BM), cl);
    const long long cost160 = static_cast<long long>(ceil_div(rows * ceil_div(N, 160), slots)) * 160;
    const long long cost128 = static_cast<long long>(ceil_div(rows * ceil_div(N, 128), slots)) * 128;
    bn = (N <= 128 || cost128 * 10 < cost160 * 9) ? 128 : 160;
  }

  CUtensorMap tmA, tmB, tmC, tmR;
  {
    const uint64_t dims[4] = {static_cast<uint64_t>(Cin), static_cast<uint64_t>(Wd), static_cast<uint64_t>(H),
                              static_cast<uint64_t>(images)};
    const uint64_t strides[3] = {static_cast<uint64_t>(Cin) * 2, static_cast<uint64_t>(Wd) * Cin * 2,
                                 static_cast<uint64_t>(H) * Wd * Cin * 2};
    const uint32_t box[4] = {GEMM_BK, static_cast<uint32_t>(OW * stride), 1, 1};
    const uint32_t estr[4] = {1, static_cast<uint32_t>(stride), 1, 1};
    int rc = make_tmap_bf16_ex(&tmA, X, 4, dims, strides, box, estr, 128);
    if (rc != FMC_OK) return rc;
  }
  {
    const uint64_t dims[2] = {static_cast<uint64_t>(K), static_cast<uint64_t>(N)};
    const uint64_t strides[1] = {static_cast<uint64_t>(K) * 2};
    const uint32_t box[2] = {GEMM_BK, static_cast<uint32_t>(use_cluster ? bn / 2 : bn)};
    int rc = make_tmap_bf16(&tmB, W, 2, dims, strides, box, true);
    if (rc != FMC_OK) return rc;
  }
  const uint32_t obox[2] = {SUB_COLS, GEMM_BM};
  {
    const uint64_t dims[2] = {static_cast<uint64_t>(N), static_cast<uint64_t>(M)};
    const uint64_t strides[1] = {static_cast<uint64_t>(N) * 2};
    int rc = make_tmap_bf16_sw(&tmC, Out, 2, dims, strides, obox, 64);
    if (rc != FMC_OK) return rc;
    if (residual != nullptr) {
      FMC_REQUIRE((reinterpret_cast<uintptr_t>(residual) & 15) == 0, FMC_ERR_SHAPE, "fmc_conv3x3_bf16: residual alignment");
      rc = make_tmap_bf16_sw(&tmR, residual, 2, dims, strides, obox, 64);
      if (rc != FMC_OK) return rc;
    } else {
      tmR = tmC;
    }
  }
  GemmParams p{};
  p.M = M; p.N = N; p.K = K;
  p.C = Out; p.ldc = N;
  p.bias = bias;
  p.residual = static_cast<const __nv_bfloat16*>(residual); p.ldr = N;
  p.rowbias = nullptr; p.rows_per_group = 1; p.ldrb = 0;
  p.flags = 0;
  p.f16_col0 = 0x7FFFFFFF;
  p.conv_oh = OH; p.conv_ow = OW; p.conv_stride = stride; p.conv_cchunks = Cin / GEMM_BK; p.conv_rows = GEMM_BM / OW;
  if (use_cluster) {
    if (bn == 128) return launch_gemm2<128, false, 2, true>(tmA, tmB, tmC, tmR, p, stream);
    return launch_gemm2<160, false, 2, true>(tmA, tmB, tmC, tmR, p, stream);
  }
  if (bn == 128) return launch_gemm2<128, false, 1, true>(tmA, tmB, tmC, tmR, p, stream);
  return launch_gemm2<160, false, 1, true>(tmA, tmB, tmC, tmR, p, stream);
}
