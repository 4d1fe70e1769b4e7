// Normalisation kernels on channels-last bf16 activations (HBM-bound; fp32 statistics).
//   LayerNorm (+ temporal positional encoding + CameraAdapter pose add)   fmc/models/motion_module.py:289,320,355-356
//                                                                          fmc/models/attention_processor.py:257 (x + pose)
//   GroupNorm(32) per frame (+ SiLU)                                        fmc/models/resnet.py:27-37, diffusers ResnetBlock2D,
//                                                                          Transformer2DModel.norm, motion_module.py:217
#include <cuda_bf16.h>

#include "common.cuh"
#include <cstdlib>

#include "ptx.cuh"

namespace fmc {

// ------------------------------------------------------------------------------------------------
// LayerNorm: one warp per row, R rows in flight per warp (all loads issued before the first reduction), the rows live
// in registers (NV vectors of 8 channels per lane, C <= NV * 256).
// out  = LN(x) * gamma + beta (+ pe[frame])            (bf16)
// out2 = out_fp32 + add                                 (bf16, optional: CameraAdapter input x + pose)
// ------------------------------------------------------------------------------------------------
constexpr int LN_MAX_VEC = 5;  // per lane: 5 vectors of 8 channels -> C <= 1280
constexpr int LN_WARPS = 8;
constexpr int LN_ADD_R_DEFAULT = 1;  // measured: 72 -> 53 us at level 0 (profiles/r01_norm_bench.txt)

template <int NV, int R, bool HAS_ADD>
__global__ void __launch_bounds__(LN_WARPS * 32)
layernorm_kernel(const __nv_bfloat16* __restrict__ x, long long ldx, const float* __restrict__ gamma,
                 const float* __restrict__ beta, float eps, __nv_bfloat16* __restrict__ out, long long ldo,
                 const float* __restrict__ pe, int F, int HW, const __nv_bfloat16* __restrict__ add, long long ldadd,
                 __nv_bfloat16* __restrict__ out2, long long ldo2, long long rows, int C) {
  pdl_launch_dependents();
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const long long row0 = (blockIdx.x * static_cast<long long>(LN_WARPS) + (threadIdx.x >> 5)) * R;
  if (row0 >= rows) return;
  const int nvec = C >> 3;
  const float inv_c = 1.0f / static_cast<float>(C);
  uint4 raw[R][NV];
#pragma unroll
  for (int r = 0; r < R; ++r) {
    const long long row = row0 + r < rows ? row0 + r : rows - 1;
    const uint4* xr = reinterpret_cast<const uint4*>(x + row * ldx);
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int vi = lane + i * 32;
      raw[r][i] = vi < nvec ? __ldg(xr + vi) : make_uint4(0u, 0u, 0u, 0u);
    }
  }
  uint4 addraw[HAS_ADD ? R : 1][HAS_ADD ? NV : 1];
  if constexpr (HAS_ADD) {
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const long long row = row0 + r < rows ? row0 + r : rows - 1;
      const uint4* ar = reinterpret_cast<const uint4*>(add + row * ldadd);
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const int vi = lane + i * 32;
        addraw[r][i] = vi < nvec ? __ldg(ar + vi) : make_uint4(0u, 0u, 0u, 0u);
      }
    }
  }
  float mean[R], rstd[R];
#pragma unroll
  for (int r = 0; r < R; ++r) {
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const uint32_t w[4] = {raw[r][i].x, raw[r][i].y, raw[r][i].z, raw[r][i].w};
#pragma unroll
      for (int j = 0; j < 4; ++j) sum += bf16_lo(w[j]) + bf16_hi(w[j]);  // padding vectors are zero
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    mean[r] = sum * inv_c;
    float sq = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      if (lane + i * 32 < nvec) {
        const uint32_t w[4] = {raw[r][i].x, raw[r][i].y, raw[r][i].z, raw[r][i].w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float d0 = bf16_lo(w[j]) - mean[r], d1 = bf16_hi(w[j]) - mean[r];
          sq += d0 * d0 + d1 * d1;
        }
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
    rstd[r] = rsqrtf(sq * inv_c + eps);
  }
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int vi = lane + i * 32;
    if (vi < nvec) {
      const int c0 = vi * 8;
      float g[8], bb[8];
#pragma unroll
      for (int j = 0; j < 8; j += 4) {
        const float4 g4 = __ldg(reinterpret_cast<const float4*>(gamma + c0 + j));
        const float4 b4 = __ldg(reinterpret_cast<const float4*>(beta + c0 + j));
        g[j] = g4.x; g[j + 1] = g4.y; g[j + 2] = g4.z; g[j + 3] = g4.w;
        bb[j] = b4.x; bb[j + 1] = b4.y; bb[j + 2] = b4.z; bb[j + 3] = b4.w;
      }
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const long long row = row0 + r;
        if (row < rows) {
          const uint32_t w[4] = {raw[r][i].x, raw[r][i].y, raw[r][i].z, raw[r][i].w};
          float y[8];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            y[2 * j] = (bf16_lo(w[j]) - mean[r]) * rstd[r] * g[2 * j] + bb[2 * j];
            y[2 * j + 1] = (bf16_hi(w[j]) - mean[r]) * rstd[r] * g[2 * j + 1] + bb[2 * j + 1];
          }
          if (pe != nullptr) {
            const float* per = pe + static_cast<long long>((row / HW) % F) * C + c0;
#pragma unroll
            for (int j = 0; j < 8; j += 4) {
              const float4 e = __ldg(reinterpret_cast<const float4*>(per + j));
              y[j] += e.x; y[j + 1] += e.y; y[j + 2] += e.z; y[j + 3] += e.w;
            }
          }
          *reinterpret_cast<uint4*>(out + row * ldo + c0) = make_uint4(
              pack_bf16x2(y[0], y[1]), pack_bf16x2(y[2], y[3]), pack_bf16x2(y[4], y[5]), pack_bf16x2(y[6], y[7]));
          if constexpr (HAS_ADD) {
            const uint32_t aw[4] = {addraw[r][i].x, addraw[r][i].y, addraw[r][i].z, addraw[r][i].w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              y[2 * j] += bf16_lo(aw[j]);
              y[2 * j + 1] += bf16_hi(aw[j]);
            }
            *reinterpret_cast<uint4*>(out2 + row * ldo2 + c0) = make_uint4(
                pack_bf16x2(y[0], y[1]), pack_bf16x2(y[2], y[3]), pack_bf16x2(y[4], y[5]), pack_bf16x2(y[6], y[7]));
          }
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// LayerNorm for C = 40 * LPR (the U-Net widths 320 / 640 / 1280): LPR lanes share a row (5 vectors of 8 channels per
// lane, every lane busy), 32 / LPR rows per warp pass, R passes in flight.  Same arithmetic as above.
// ------------------------------------------------------------------------------------------------
template <int LPR, int R, bool HAS_ADD>
__global__ void __launch_bounds__(LN_WARPS * 32)
layernorm40_kernel(const __nv_bfloat16* __restrict__ x, long long ldx, const float* __restrict__ gamma,
                   const float* __restrict__ beta, float eps, __nv_bfloat16* __restrict__ out, long long ldo,
                   const float* __restrict__ pe, int F, int HW, const __nv_bfloat16* __restrict__ add, long long ldadd,
                   __nv_bfloat16* __restrict__ out2, long long ldo2, long long rows) {
  pdl_launch_dependents();
  pdl_wait();
  constexpr int C = 40 * LPR;
  constexpr int RPW = 32 / LPR;  // rows per warp pass
  constexpr int NV = 5;
  const int lane = threadIdx.x & 31;
  const int sub = lane % LPR;    // lane inside the row group
  const int rw = lane / LPR;     // row inside the warp pass
  const long long row0 = (blockIdx.x * static_cast<long long>(LN_WARPS) + (threadIdx.x >> 5)) * (RPW * R) + rw;
  if (row0 - rw >= rows) return;
  constexpr float inv_c = 1.0f / static_cast<float>(C);
  uint4 raw[R][NV];
#pragma unroll
  for (int r = 0; r < R; ++r) {
    const long long row = row0 + r * RPW < rows ? row0 + r * RPW : rows - 1;
    const uint4* xr = reinterpret_cast<const uint4*>(x + row * ldx);
#pragma unroll
    for (int i = 0; i < NV; ++i) raw[r][i] = __ldg(xr + sub + i * LPR);
  }
  uint4 addraw[HAS_ADD ? R : 1][HAS_ADD ? NV : 1];
  if constexpr (HAS_ADD) {
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const long long row = row0 + r * RPW < rows ? row0 + r * RPW : rows - 1;
      const uint4* ar = reinterpret_cast<const uint4*>(add + row * ldadd);
#pragma unroll
      for (int i = 0; i < NV; ++i) addraw[r][i] = __ldg(ar + sub + i * LPR);
    }
  }
  float mean[R], rstd[R];
#pragma unroll
  for (int r = 0; r < R; ++r) {
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const uint32_t w[4] = {raw[r][i].x, raw[r][i].y, raw[r][i].z, raw[r][i].w};
#pragma unroll
      for (int j = 0; j < 4; ++j) sum += bf16_lo(w[j]) + bf16_hi(w[j]);
    }
#pragma unroll
    for (int o = LPR / 2; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    mean[r] = sum * inv_c;
    float sq = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const uint32_t w[4] = {raw[r][i].x, raw[r][i].y, raw[r][i].z, raw[r][i].w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float d0 = bf16_lo(w[j]) - mean[r], d1 = bf16_hi(w[j]) - mean[r];
        sq += d0 * d0 + d1 * d1;
      }
    }
#pragma unroll
    for (int o = LPR / 2; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
    rstd[r] = rsqrtf(sq * inv_c + eps);
  }
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c0 = (sub + i * LPR) * 8;
    float g[8], bb[8];
#pragma unroll
    for (int j = 0; j < 8; j += 4) {
      const float4 g4 = __ldg(reinterpret_cast<const float4*>(gamma + c0 + j));
      const float4 b4 = __ldg(reinterpret_cast<const float4*>(beta + c0 + j));
      g[j] = g4.x; g[j + 1] = g4.y; g[j + 2] = g4.z; g[j + 3] = g4.w;
      bb[j] = b4.x; bb[j + 1] = b4.y; bb[j + 2] = b4.z; bb[j + 3] = b4.w;
    }
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const long long row = row0 + r * RPW;
      if (row < rows) {
        const uint32_t w[4] = {raw[r][i].x, raw[r][i].y, raw[r][i].z, raw[r][i].w};
        float y[8];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          y[2 * j] = (bf16_lo(w[j]) - mean[r]) * rstd[r] * g[2 * j] + bb[2 * j];
          y[2 * j + 1] = (bf16_hi(w[j]) - mean[r]) * rstd[r] * g[2 * j + 1] + bb[2 * j + 1];
        }
        if (pe != nullptr) {
          const float* per = pe + static_cast<long long>((row / HW) % F) * C + c0;
#pragma unroll
          for (int j = 0; j < 8; j += 4) {
            const float4 e = __ldg(reinterpret_cast<const float4*>(per + j));
            y[j] += e.x; y[j + 1] += e.y; y[j + 2] += e.z; y[j + 3] += e.w;
          }
        }
        *reinterpret_cast<uint4*>(out + row * ldo + c0) = make_uint4(
            pack_bf16x2(y[0], y[1]), pack_bf16x2(y[2], y[3]), pack_bf16x2(y[4], y[5]), pack_bf16x2(y[6], y[7]));
        if constexpr (HAS_ADD) {
          const uint32_t aw[4] = {addraw[r][i].x, addraw[r][i].y, addraw[r][i].z, addraw[r][i].w};
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            y[2 * j] += bf16_lo(aw[j]);
            y[2 * j + 1] += bf16_hi(aw[j]);
          }
          *reinterpret_cast<uint4*>(out2 + row * ldo2 + c0) = make_uint4(
              pack_bf16x2(y[0], y[1]), pack_bf16x2(y[2], y[3]), pack_bf16x2(y[4], y[5]), pack_bf16x2(y[6], y[7]));
        }
      }
    }
  }
}

template <int LPR, int R>
static int launch_layernorm40(const void* x, long long ldx, const float* gamma, const float* beta, float eps, void* out,
                              long long ldo, const float* pe, int F, int HW, const void* add, long long ldadd,
                              void* out2, long long ldo2, long long rows, cudaStream_t stream) {
  const long long per_block = static_cast<long long>(LN_WARPS) * (32 / LPR) * R;
  const unsigned grid = static_cast<unsigned>((rows + per_block - 1) / per_block);
  if (out2 != nullptr) {
    launch_k(layernorm40_kernel<LPR, R, true>, dim3(grid), dim3(LN_WARPS * 32), 0, stream, 
        static_cast<const __nv_bfloat16*>(x), ldx, gamma, beta, eps, static_cast<__nv_bfloat16*>(out), ldo, pe,
        F > 0 ? F : 1, HW > 0 ? HW : 1, static_cast<const __nv_bfloat16*>(add), ldadd,
        static_cast<__nv_bfloat16*>(out2), ldo2, rows);
  } else {
    launch_k(layernorm40_kernel<LPR, R, false>, dim3(grid), dim3(LN_WARPS * 32), 0, stream, 
        static_cast<const __nv_bfloat16*>(x), ldx, gamma, beta, eps, static_cast<__nv_bfloat16*>(out), ldo, pe,
        F > 0 ? F : 1, HW > 0 ? HW : 1, nullptr, 0, nullptr, 0, rows);
  }
  return check_launch("layernorm40_kernel");
}

// ------------------------------------------------------------------------------------------------
// (Folding the finalize step into the apply kernel -- every apply block reducing its image's partial sums itself, one
// launch fewer per GroupNorm -- was measured 1-2 % SLOWER on the whole step in round 1: ~1700 apply blocks per level-0
// call each pay the extra dependent L2 round trip before they start streaming.)
// GroupNorm, deterministic two-kernel form (no atomics: a fixed reduction order, so two runs are bit-identical).
//   partial: grid (chunks, images); a block reduces GN_ROWS rows of one image to (sum, sum of squares) per group and
//            writes partial[img][chunk][g][2].
//   apply  : grid (chunks, images); a block folds the partials of its image in chunk order into mean / rstd, turns
//            gamma / beta (and the optional per-image channel bias) into one scale / shift pair per channel, then
//            streams its rows: y = x * a + b (optional SiLU).
// Thread layout in both: `nvec = C / 8` lanes per row (one 16-byte vector each), blockDim / nvec rows in flight.
// ------------------------------------------------------------------------------------------------
// y * sigmoid(y) on a packed pair: the same arithmetic as the scalar `y * rcp_fast(1 + __expf(-y))` of the apply kernel
// (ex2 of y * -log2(e), + 1, MUFU.RCP, * y) with the three fp32 operations issued as .f32x2 and the exponential without
// the denormal-range fix-up of the non-ftz form (an exponent below -126 flushes to 0, and 1 + 0 == 1 + denormal)
__device__ __forceinline__ uint64_t silu_f2(uint64_t y2) {
  float t0, t1;
  f2_unpack(f2_mul(y2, f2_pack(-1.4426950408889634f, -1.4426950408889634f)), t0, t1);
  float e0, e1;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"(t0));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(t1));
  float d0, d1;
  f2_unpack(f2_add(f2_pack(e0, e1), f2_pack(1.0f, 1.0f)), d0, d1);
  return f2_mul(y2, f2_pack(rcp_fast(d0), rcp_fast(d1)));
}
__device__ __forceinline__ uint64_t bf16x2_to_f2(uint32_t w) { return f2_pack(bf16_lo(w), bf16_hi(w)); }
__device__ __forceinline__ uint32_t f2_to_bf16x2(uint64_t v) {
  float lo, hi;
  f2_unpack(v, lo, hi);
  return pack_bf16x2(lo, hi);
}

constexpr int GN_ROWS = 64;       // rows per block, partial kernel
constexpr int GN_APPLY_ROWS = 32;  // rows per block, apply kernel
constexpr int GN_MAX_GROUPS = 64;

__global__ void __launch_bounds__(1024)
groupnorm_partial_kernel(const __nv_bfloat16* __restrict__ x, long long ldx, float* __restrict__ partial, int HW, int C,
                         int G, int chunks, const float* __restrict__ rowbias, long long ldrb, int rb_div) {
  pdl_launch_dependents();
  pdl_wait();
  extern __shared__ float gn_smem[];  // [rows_par][C] sums, then [rows_par][C] squares
  const int img = blockIdx.y;
  const int row0 = blockIdx.x * GN_ROWS;
  const int nvec = C >> 3;
  const int rows_par = blockDim.x / nvec;
  const int vi = threadIdx.x % nvec;
  const int rsub = threadIdx.x / nvec;
  float a[8], b[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) a[j] = b[j] = 0.f;
  if (rsub < rows_par) {
    uint64_t rb2[4], a2[4], b2[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      rb2[j] = rowbias != nullptr ? f2_pack(__ldg(rowbias + (img / rb_div) * ldrb + vi * 8 + 2 * j),
                                            __ldg(rowbias + (img / rb_div) * ldrb + vi * 8 + 2 * j + 1)) : 0ull;
      a2[j] = b2[j] = 0ull;
    }
    const int rend = min(row0 + GN_ROWS, HW);
    for (int r = row0 + rsub; r < rend; r += rows_par) {
      const uint4 u = __ldg(reinterpret_cast<const uint4*>(x + (static_cast<long long>(img) * HW + r) * ldx) + vi);
      const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const uint64_t xr = f2_add(bf16x2_to_f2(w[j]), rb2[j]);
        a2[j] = f2_add(a2[j], xr);
        b2[j] = f2_fma(xr, xr, b2[j]);
      }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      f2_unpack(a2[j], a[2 * j], a[2 * j + 1]);
      f2_unpack(b2[j], b[2 * j], b[2 * j + 1]);
    }
    float* ssum = gn_smem + static_cast<size_t>(rsub) * C + vi * 8;
    float* ssq = gn_smem + static_cast<size_t>(rows_par + rsub) * C + vi * 8;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      ssum[j] = a[j];
      ssq[j] = b[j];
    }
  }
  __syncthreads();
  // fold rows first (2*C column sums over rows_par rows), then channels into groups -- both in a fixed order
  for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) {
    const int which = i / C, c = i - which * C;
    float acc = 0.f;
    for (int r = 0; r < rows_par; ++r) acc += gn_smem[static_cast<size_t>(which * rows_par + r) * C + c];
    gn_smem[static_cast<size_t>(which * rows_par) * C + c] = acc;  // row 0 of each half now holds the column sums
  }
  __syncthreads();
  if (threadIdx.x < 2 * G) {
    const int g = threadIdx.x >> 1, which = threadIdx.x & 1;
    const int cpg = C / G;
    const float* src = gn_smem + static_cast<size_t>(which * rows_par) * C + g * cpg;
    float acc = 0.f;
    for (int c = 0; c < cpg; ++c) acc += src[c];
    partial[((static_cast<long long>(img) * chunks + blockIdx.x) * G + g) * 2 + which] = acc;
  }
}

// scale / shift per (image, channel): y = x * a + b with a = rstd * gamma, b = beta + (rowbias - mean) * a.
// grid = images; the partial sums are folded in chunk order (deterministic).
__global__ void __launch_bounds__(256)
groupnorm_finalize_kernel(const float* __restrict__ partial, const float* __restrict__ gamma,
                          const float* __restrict__ beta, float eps, float2* __restrict__ ab, int HW, int C, int G,
                          int chunks, const float* __restrict__ rowbias, long long ldrb, int rb_div) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ float s_mean[GN_MAX_GROUPS], s_rstd[GN_MAX_GROUPS];
  const int img = blockIdx.x;
  const int cpg = C / G;
  if (threadIdx.x < G) {
    float s = 0.f, ss = 0.f;
    const float* pp = partial + (static_cast<long long>(img) * chunks * G + threadIdx.x) * 2;
    for (int c = 0; c < chunks; ++c) {
      s += pp[static_cast<long long>(c) * G * 2];
      ss += pp[static_cast<long long>(c) * G * 2 + 1];
    }
    const float inv_n = 1.0f / (static_cast<float>(HW) * cpg);
    const float mean = s * inv_n;
    const float var = fmaxf(ss * inv_n - mean * mean, 0.f);
    s_mean[threadIdx.x] = mean;
    s_rstd[threadIdx.x] = rsqrtf(var + eps);
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const int g = c / cpg;
    const float a = s_rstd[g] * __ldg(gamma + c);
    const float rbv = rowbias != nullptr ? __ldg(rowbias + (img / rb_div) * ldrb + c) : 0.f;
    ab[static_cast<long long>(img) * C + c] = make_float2(a, fmaf(rbv - s_mean[g], a, __ldg(beta + c)));
  }
}

constexpr int GN_UNROLL = 4;

__global__ void __launch_bounds__(1024)
groupnorm_apply_kernel(const __nv_bfloat16* __restrict__ x, long long ldx, const float2* __restrict__ ab,
                       __nv_bfloat16* __restrict__ out, long long ldo, int HW, int C, int silu) {
  pdl_launch_dependents();
  pdl_wait();
  const int img = blockIdx.y;
  const int nvec = C >> 3;
  const int rows_par = blockDim.x / nvec;
  const int vi = threadIdx.x % nvec;
  const int rsub = threadIdx.x / nvec;
  if (rsub >= rows_par) return;
  uint64_t sa2[4], sb2[4];
  {
    const float4* p4 = reinterpret_cast<const float4*>(ab + static_cast<long long>(img) * C + vi * 8);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float4 v = __ldg(p4 + j);
      sa2[j] = f2_pack(v.x, v.z);
      sb2[j] = f2_pack(v.y, v.w);
    }
  }
  const int row0 = blockIdx.x * (rows_par * GN_UNROLL) + rsub;
  uint4 u[GN_UNROLL];
#pragma unroll
  for (int k = 0; k < GN_UNROLL; ++k) {
    const int r = row0 + k * rows_par;
    if (r < HW) u[k] = __ldg(reinterpret_cast<const uint4*>(x + (static_cast<long long>(img) * HW + r) * ldx) + vi);
  }
#pragma unroll
  for (int k = 0; k < GN_UNROLL; ++k) {
    const int r = row0 + k * rows_par;
    if (r < HW) {
      const uint32_t w[4] = {u[k].x, u[k].y, u[k].z, u[k].w};
      uint32_t o[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        uint64_t y2 = f2_fma(bf16x2_to_f2(w[j]), sa2[j], sb2[j]);
        if (silu) y2 = silu_f2(y2);
        o[j] = f2_to_bf16x2(y2);
      }
      *reinterpret_cast<uint4*>(out + (static_cast<long long>(img) * HW + r) * ldo + vi * 8) =
          make_uint4(o[0], o[1], o[2], o[3]);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// GroupNorm in ONE pass over HBM (read x once, write y once; the three-kernel form above reads x twice and pays three
// dependent launches, which is all that is left of a GroupNorm at the two coarse U-Net levels).
//   work item = (image, channel chunk): CH = lcm(C / G, 8) channels, i.e. whole groups AND whole 16-byte vectors
//               (40 channels = 4 groups at C = 320, 2 groups at C = 640, 1 at C = 1280; 80 at C = 2560; 120 at
//               C = 960 / 1920).  A cluster of R CTAs (R = 1, 2, 4, 8) splits the HW rows of the item; each CTA keeps
//               its rows x CH slab in REGISTERS (<= VMAX 16-byte vectors per thread), reduces it to per-group
//               (sum, sum of squares), publishes the pair in its shared memory, and after one cluster barrier every
//               CTA folds the R partials in rank order through distributed shared memory -- a fixed order, so all CTAs
//               get the same bits and two runs are identical.  Then y = x * a + b (+ SiLU) straight from the registers.
// Thread layout: VPR = CH / 8 lanes per row, 256 / VPR rows per pass, like the kernels above.
// ------------------------------------------------------------------------------------------------
constexpr int GN_FUSED_DEFAULT = 4;  // measured (profiles/r01_norm_bench.txt): GroupNorm 3.33 (mode 0) -> 2.43 (1) -> 2.21 ms (3) per step
constexpr int GNF_THREADS = 256;
constexpr int GNF_MAX_CH = 128;     // VPR <= 16
constexpr int GNF_MAX_CLUSTER = 8;  // portable cluster size

__device__ __forceinline__ float2 ld_dsmem_f2(uint32_t cluster_addr) {
  float2 v;
  asm volatile("ld.shared::cluster.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(cluster_addr) : "memory");
  return v;
}
__device__ __forceinline__ void cluster_arrive_release() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
}
__device__ __forceinline__ void cluster_wait_acquire() {
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

template <int VMAX>
__global__ void __launch_bounds__(GNF_THREADS)
groupnorm_fused_kernel(const __nv_bfloat16* __restrict__ x, long long ldx, const float* __restrict__ gamma,
                       const float* __restrict__ beta, float eps, __nv_bfloat16* __restrict__ out, long long ldo,
                       int HW, int cpg, int CH, int rows_cta, int silu, const float* __restrict__ rowbias,
                       long long ldrb, int rb_div) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ float red_s[2 * GNF_THREADS], red_q[2 * GNF_THREADS];  // [slot = vc * 2 + which][rsub]
  __shared__ float tot_s[2 * GNF_MAX_CH / 8], tot_q[2 * GNF_MAX_CH / 8];
  __shared__ float2 part[8];                                         // this CTA's (sum, sum of squares) per group
  __shared__ float s_mean[8], s_rstd[8];
  const int tid = threadIdx.x;
  const int R = gridDim.x, rank = blockIdx.x, chunk = blockIdx.y, img = blockIdx.z;
  const int VPR = CH >> 3;
  const int rows_par = GNF_THREADS / VPR;
  const int vc = tid % VPR, rsub = tid / VPR;
  const bool active = rsub < rows_par;
  const int ngc = CH / cpg;                       // groups per chunk (<= 8)
  const int c0 = chunk * CH + vc * 8;             // first channel of this thread's vector
  const int g_first = c0 / cpg;                   // group of channel c0 (global index)
  const int split = min(8, (g_first + 1) * cpg - c0);  // channels [0, split) -> g_first, [split, 8) -> g_first + 1
  const int row_begin = rank * rows_cta;
  const int row_end = min(HW, row_begin + rows_cta);
  // this thread's rows: row_begin + rsub + i * rows_par for i < nvalid
  const int mine = row_end - row_begin - rsub;
  const int nvalid = (active && mine > 0) ? min(VMAX, (mine + rows_par - 1) / rows_par) : 0;
  const long long first_row = static_cast<long long>(img) * HW + row_begin + rsub;
  const uint4* xp = reinterpret_cast<const uint4*>(x + first_row * ldx + c0);
  const long long xstep = static_cast<long long>(rows_par) * ldx / 8;  // in 16-byte vectors (ldx % 8 == 0)

  uint4 v[VMAX];
#pragma unroll
  for (int i = 0; i < VMAX; ++i) v[i] = i < nvalid ? __ldg(xp + i * xstep) : make_uint4(0u, 0u, 0u, 0u);
  uint64_t rb2[4];
  if (rowbias != nullptr) {
    const float4* rp = reinterpret_cast<const float4*>(rowbias + (img / rb_div) * ldrb + c0);
    const float4 r0 = __ldg(rp), r1 = __ldg(rp + 1);
    rb2[0] = f2_pack(r0.x, r0.y); rb2[1] = f2_pack(r0.z, r0.w); rb2[2] = f2_pack(r1.x, r1.y); rb2[3] = f2_pack(r1.z, r1.w);
  } else {
#pragma unroll
    for (int j = 0; j < 4; ++j) rb2[j] = 0ull;
  }

  // ---- per-thread channel sums (packed fp32 pairs), split into the (at most two) groups the vector touches
  {
    uint64_t a2[4], b2[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) a2[j] = b2[j] = 0ull;
#pragma unroll
    for (int i = 0; i < VMAX; ++i) {
      if (i < nvalid) {
        const uint32_t w[4] = {v[i].x, v[i].y, v[i].z, v[i].w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const uint64_t xr = f2_add(bf16x2_to_f2(w[j]), rb2[j]);
          a2[j] = f2_add(a2[j], xr);
          b2[j] = f2_fma(xr, xr, b2[j]);
        }
      }
    }
    float a[8], b[8];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      f2_unpack(a2[j], a[2 * j], a[2 * j + 1]);
      f2_unpack(b2[j], b[2 * j], b[2 * j + 1]);
    }
    float s_lo = 0.f, q_lo = 0.f, s_hi = 0.f, q_hi = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      if (j < split) { s_lo += a[j]; q_lo += b[j]; } else { s_hi += a[j]; q_hi += b[j]; }
    }
    if (active) {
      red_s[(vc * 2) * rows_par + rsub] = s_lo;
      red_q[(vc * 2) * rows_par + rsub] = q_lo;
      red_s[(vc * 2 + 1) * rows_par + rsub] = s_hi;
      red_q[(vc * 2 + 1) * rows_par + rsub] = q_hi;
    }
  }
  __syncthreads();
  // ---- rows -> one pair per slot (a warp per slot, fixed butterfly), slots -> groups (slot order)
  {
    const int warp = tid >> 5, lane = tid & 31;
    for (int s = warp; s < 2 * VPR; s += GNF_THREADS / 32) {
      float ss = 0.f, qq = 0.f;
      for (int r = lane; r < rows_par; r += 32) {
        ss += red_s[s * rows_par + r];
        qq += red_q[s * rows_par + r];
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        ss += __shfl_xor_sync(0xffffffffu, ss, o);
        qq += __shfl_xor_sync(0xffffffffu, qq, o);
      }
      if (lane == 0) { tot_s[s] = ss; tot_q[s] = qq; }
    }
  }
  __syncthreads();
  if (tid < ngc) {
    const int g = chunk * ngc + tid;
    float ss = 0.f, qq = 0.f;
    for (int s = 0; s < 2 * VPR; ++s) {
      const int sc0 = chunk * CH + (s >> 1) * 8;
      if (sc0 / cpg + (s & 1) == g) { ss += tot_s[s]; qq += tot_q[s]; }
    }
    part[tid] = make_float2(ss, qq);
  }
  // ---- cluster exchange: every CTA folds the R partials in rank order
  if (R > 1) { cluster_arrive_release(); cluster_wait_acquire(); } else { __syncthreads(); }
  if (tid < ngc) {
    float ss = 0.f, qq = 0.f;
    if (R > 1) {
      // all remote loads first (independent, ~0.5 us each when serialised behind their adds), then the fixed-order fold
      const uint32_t mine_addr = smem_u32(&part[tid]);
      float2 p[GNF_MAX_CLUSTER];
#pragma unroll
      for (int r = 0; r < GNF_MAX_CLUSTER; ++r)
        p[r] = r < R ? ld_dsmem_f2(mapa_shared(mine_addr, static_cast<uint32_t>(r))) : make_float2(0.f, 0.f);
#pragma unroll
      for (int r = 0; r < GNF_MAX_CLUSTER; ++r) {
        if (r < R) { ss += p[r].x; qq += p[r].y; }
      }
    } else {
      ss = part[tid].x; qq = part[tid].y;
    }
    const float inv_n = 1.0f / (static_cast<float>(HW) * cpg);
    const float mean = ss * inv_n;
    const float var = fmaxf(qq * inv_n - mean * mean, 0.f);
    s_mean[tid] = mean;
    s_rstd[tid] = rsqrtf(var + eps);
  }
  __syncthreads();
  if (R > 1) cluster_arrive_release();  // my remote reads are done; peers may exit once everybody got here

  // ---- apply from registers: y = x * a + b with a = rstd * gamma, b = beta + (rowbias - mean) * a
  if (nvalid > 0) {
    float sa[8], sb[8], rbf[8];
#pragma unroll
    for (int j = 0; j < 8; j += 4) {
      const float4 g4 = __ldg(reinterpret_cast<const float4*>(gamma + c0 + j));
      const float4 b4 = __ldg(reinterpret_cast<const float4*>(beta + c0 + j));
      sa[j] = g4.x; sa[j + 1] = g4.y; sa[j + 2] = g4.z; sa[j + 3] = g4.w;
      sb[j] = b4.x; sb[j + 1] = b4.y; sb[j + 2] = b4.z; sb[j + 3] = b4.w;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) f2_unpack(rb2[j], rbf[2 * j], rbf[2 * j + 1]);
    const int gl = g_first - chunk * ngc;
    const float m0 = s_mean[gl], r0 = s_rstd[gl];
    const float m1 = split < 8 ? s_mean[gl + 1] : 0.f, r1 = split < 8 ? s_rstd[gl + 1] : 0.f;
    uint64_t sa2[4], sb2[4];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float a = (j < split ? r0 : r1) * sa[j];
      sb[j] = fmaf(rbf[j] - (j < split ? m0 : m1), a, sb[j]);
      sa[j] = a;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      sa2[j] = f2_pack(sa[2 * j], sa[2 * j + 1]);
      sb2[j] = f2_pack(sb[2 * j], sb[2 * j + 1]);
    }
    uint4* op = reinterpret_cast<uint4*>(out + first_row * ldo + c0);
    const long long ostep = static_cast<long long>(rows_par) * ldo / 8;
#pragma unroll
    for (int i = 0; i < VMAX; ++i) {
      if (i < nvalid) {
        const uint32_t w[4] = {v[i].x, v[i].y, v[i].z, v[i].w};
        uint32_t o[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          uint64_t y2 = f2_fma(bf16x2_to_f2(w[j]), sa2[j], sb2[j]);
          if (silu) y2 = silu_f2(y2);
          o[j] = f2_to_bf16x2(y2);
        }
        op[i * ostep] = make_uint4(o[0], o[1], o[2], o[3]);
      }
    }
  }
  if (R > 1) cluster_wait_acquire();
}

// ------------------------------------------------------------------------------------------------
// The same single-pass GroupNorm with the slab staged in SHARED memory by TMA instead of registers.  Why: the register
// kernel above runs 3 CTAs per SM (77 registers) and a CTA has loads in flight only for the first third of its life
// (load -> reduce -> cluster barrier -> apply -> store), so an SM averages ~32 KB in flight and the kernel sits at
// 2.3 TB/s whether x comes from HBM or L2 (profiles/r01_norm_bench.txt: `hot` == `cold`).  Here one elected thread
// issues the whole rows_cta x CH slab as 3-D TMA boxes (channel chunk, rows, image; rows past the image are
// zero-filled and masked) the moment the CTA starts, nothing is held in registers across the barriers, and 5 CTAs fit
// an SM (<= 32 KB slab + 4.5 KB reduction scratch each, 48 registers; 6 CTAs = 40 registers spill).
// ------------------------------------------------------------------------------------------------
constexpr int GNT_MAX_SLAB = 32768;
constexpr int GNT_MIN_SLAB = 16384;  // below this the TMA round trip costs more than direct loads (measured: 12.8 KB slabs lose 10 %)

__global__ void __launch_bounds__(GNF_THREADS, 5)
groupnorm_tma_kernel(const __grid_constant__ CUtensorMap tmx, const float* __restrict__ gamma,
                     const float* __restrict__ beta, float eps, __nv_bfloat16* __restrict__ out, long long ldo, int HW,
                     int cpg, int CH, int rows_cta, int boxr, int nbox, int silu, const float* __restrict__ rowbias,
                     long long ldrb, int rb_div) {
  pdl_launch_dependents();
  extern __shared__ uint8_t gnt_raw[];
  __shared__ uint64_t bar;
  __shared__ float red_s[2 * GNF_THREADS], red_q[2 * GNF_THREADS];
  __shared__ float tot_s[2 * GNF_MAX_CH / 8], tot_q[2 * GNF_MAX_CH / 8];
  __shared__ float2 part[8];
  __shared__ float s_mean[8], s_rstd[8];
  const int tid = threadIdx.x;
  const int R = gridDim.x, rank = blockIdx.x, chunk = blockIdx.y, img = blockIdx.z;
  const int row_begin = rank * rows_cta;
  const int nrows = min(HW, row_begin + rows_cta) - row_begin;  // > 0 by construction of the plan
  const uint32_t slab = (smem_u32(gnt_raw) + 127u) & ~127u;
  if (tid == 0) {
    mbar_init(&bar, 1);
    fence_barrier_init();
  }
  __syncthreads();
  pdl_wait();
  if (tid < 32) {
    if (elect_one()) {
      mbar_arrive_expect_tx(&bar, static_cast<uint32_t>(nbox * boxr * CH * 2));
      for (int b = 0; b < nbox; ++b) {
        asm volatile(
            "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
            ::"r"(slab + static_cast<uint32_t>(b * boxr * CH * 2)), "l"(reinterpret_cast<uint64_t>(&tmx)),
            "r"(smem_u32(&bar)), "r"(chunk * CH), "r"(row_begin + b * boxr), "r"(img)
            : "memory");
      }
    }
  }
  const int VPR = CH >> 3;
  const int rows_par = GNF_THREADS / VPR;
  const int vc = tid % VPR, rsub = tid / VPR;
  const bool active = rsub < rows_par;
  const int ngc = CH / cpg;
  const int c0 = chunk * CH + vc * 8;
  const int g_first = c0 / cpg;
  const int split = min(8, (g_first + 1) * cpg - c0);
  const int mine = nrows - rsub;
  const int nvalid = (active && mine > 0) ? (mine + rows_par - 1) / rows_par : 0;
  const uint32_t my0 = slab + static_cast<uint32_t>((rsub * VPR + vc) * 16);  // == slab + tid * 16 for active threads
  const uint32_t sstep = static_cast<uint32_t>(rows_par * VPR * 16);
  uint64_t rb2[4];
  if (rowbias != nullptr) {
    const float4* rp = reinterpret_cast<const float4*>(rowbias + (img / rb_div) * ldrb + c0);
    const float4 r0 = __ldg(rp), r1 = __ldg(rp + 1);
    rb2[0] = f2_pack(r0.x, r0.y); rb2[1] = f2_pack(r0.z, r0.w); rb2[2] = f2_pack(r1.x, r1.y); rb2[3] = f2_pack(r1.z, r1.w);
  } else {
#pragma unroll
    for (int j = 0; j < 4; ++j) rb2[j] = 0ull;
  }
  mbar_wait(&bar, 0);

  // ---- per-thread channel sums from the slab
  {
    uint64_t a2[4], b2[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) a2[j] = b2[j] = 0ull;
#pragma unroll 4
    for (int i = 0; i < nvalid; ++i) {
      uint32_t w[4];
      asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];"
                   : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]) : "r"(my0 + i * sstep));
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const uint64_t xr = f2_add(bf16x2_to_f2(w[j]), rb2[j]);
        a2[j] = f2_add(a2[j], xr);
        b2[j] = f2_fma(xr, xr, b2[j]);
      }
    }
    float a[8], b[8];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      f2_unpack(a2[j], a[2 * j], a[2 * j + 1]);
      f2_unpack(b2[j], b[2 * j], b[2 * j + 1]);
    }
    float s_lo = 0.f, q_lo = 0.f, s_hi = 0.f, q_hi = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      if (j < split) { s_lo += a[j]; q_lo += b[j]; } else { s_hi += a[j]; q_hi += b[j]; }
    }
    if (active) {
      red_s[(vc * 2) * rows_par + rsub] = s_lo;
      red_q[(vc * 2) * rows_par + rsub] = q_lo;
      red_s[(vc * 2 + 1) * rows_par + rsub] = s_hi;
      red_q[(vc * 2 + 1) * rows_par + rsub] = q_hi;
    }
  }
  __syncthreads();
  {
    const int warp = tid >> 5, lane = tid & 31;
    for (int s = warp; s < 2 * VPR; s += GNF_THREADS / 32) {
      float ss = 0.f, qq = 0.f;
      for (int r = lane; r < rows_par; r += 32) {
        ss += red_s[s * rows_par + r];
        qq += red_q[s * rows_par + r];
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        ss += __shfl_xor_sync(0xffffffffu, ss, o);
        qq += __shfl_xor_sync(0xffffffffu, qq, o);
      }
      if (lane == 0) { tot_s[s] = ss; tot_q[s] = qq; }
    }
  }
  __syncthreads();
  if (tid < ngc) {
    const int g = chunk * ngc + tid;
    float ss = 0.f, qq = 0.f;
    for (int s = 0; s < 2 * VPR; ++s) {
      const int sc0 = chunk * CH + (s >> 1) * 8;
      if (sc0 / cpg + (s & 1) == g) { ss += tot_s[s]; qq += tot_q[s]; }
    }
    part[tid] = make_float2(ss, qq);
  }
  if (R > 1) { cluster_arrive_release(); cluster_wait_acquire(); } else { __syncthreads(); }
  if (tid < ngc) {
    float ss = 0.f, qq = 0.f;
    if (R > 1) {
      const uint32_t mine_addr = smem_u32(&part[tid]);
      float2 p[GNF_MAX_CLUSTER];
#pragma unroll
      for (int r = 0; r < GNF_MAX_CLUSTER; ++r)
        p[r] = r < R ? ld_dsmem_f2(mapa_shared(mine_addr, static_cast<uint32_t>(r))) : make_float2(0.f, 0.f);
#pragma unroll
      for (int r = 0; r < GNF_MAX_CLUSTER; ++r) {
        if (r < R) { ss += p[r].x; qq += p[r].y; }
      }
    } else {
      ss = part[tid].x; qq = part[tid].y;
    }
    const float inv_n = 1.0f / (static_cast<float>(HW) * cpg);
    const float mean = ss * inv_n;
    const float var = fmaxf(qq * inv_n - mean * mean, 0.f);
    s_mean[tid] = mean;
    s_rstd[tid] = rsqrtf(var + eps);
  }
  __syncthreads();
  if (R > 1) cluster_arrive_release();

  // ---- apply from the slab
  if (nvalid > 0) {
    float sa[8], sb[8], rbf[8];
#pragma unroll
    for (int j = 0; j < 8; j += 4) {
      const float4 g4 = __ldg(reinterpret_cast<const float4*>(gamma + c0 + j));
      const float4 b4 = __ldg(reinterpret_cast<const float4*>(beta + c0 + j));
      sa[j] = g4.x; sa[j + 1] = g4.y; sa[j + 2] = g4.z; sa[j + 3] = g4.w;
      sb[j] = b4.x; sb[j + 1] = b4.y; sb[j + 2] = b4.z; sb[j + 3] = b4.w;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) f2_unpack(rb2[j], rbf[2 * j], rbf[2 * j + 1]);
    const int gl = g_first - chunk * ngc;
    const float m0 = s_mean[gl], r0 = s_rstd[gl];
    const float m1 = split < 8 ? s_mean[gl + 1] : 0.f, r1 = split < 8 ? s_rstd[gl + 1] : 0.f;
    uint64_t sa2[4], sb2[4];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float a = (j < split ? r0 : r1) * sa[j];
      sb[j] = fmaf(rbf[j] - (j < split ? m0 : m1), a, sb[j]);
      sa[j] = a;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      sa2[j] = f2_pack(sa[2 * j], sa[2 * j + 1]);
      sb2[j] = f2_pack(sb[2 * j], sb[2 * j + 1]);
    }
    uint4* op = reinterpret_cast<uint4*>(out + (static_cast<long long>(img) * HW + row_begin + rsub) * ldo + c0);
    const long long ostep = static_cast<long long>(rows_par) * ldo / 8;
#pragma unroll 4
    for (int i = 0; i < nvalid; ++i) {
      uint32_t w[4], o[4];
      asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];"
                   : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]) : "r"(my0 + i * sstep));
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        uint64_t y2 = f2_fma(bf16x2_to_f2(w[j]), sa2[j], sb2[j]);
        if (silu) y2 = silu_f2(y2);
        o[j] = f2_to_bf16x2(y2);
      }
      op[i * ostep] = make_uint4(o[0], o[1], o[2], o[3]);
    }
  }
  if (R > 1) cluster_wait_acquire();
}

// ------------------------------------------------------------------------------------------------
// PERSISTENT form of the TMA-staged kernel (round 2).  The kernel above is bound by bytes in flight: a CTA has loads
// outstanding only while it waits for its slab, then reduces / exchanges / applies with nothing in flight, and the CTAs
// of a wave run those phases in lockstep.  Here a cluster is resident for the whole launch and walks work items
// (image, channel chunk) with TWO slab buffers: the TMA load of item i+1 is issued before the reduction of item i
// starts, so every CTA always has a slab in flight while it computes.  Arithmetic and reduction order are those of
// groupnorm_tma_kernel -- the two produce identical bits (tests/test_gpu_ops.py::test_groupnorm_persistent_is_bit_identical).
// grid = (R, n_clusters), cluster (R, 1, 1); item = blockIdx.y + it * gridDim.y over chunks * images items.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(GNF_THREADS, 3)
groupnorm_tma_persist_kernel(const __grid_constant__ CUtensorMap tmx, const float* __restrict__ gamma,
                             const float* __restrict__ beta, float eps, __nv_bfloat16* __restrict__ out, long long ldo,
                             int HW, int cpg, int CH, int rows_cta, int boxr, int nbox, int silu,
                             const float* __restrict__ rowbias, long long ldrb, int rb_div, int chunks, int items,
                             int slab_bytes) {
  pdl_launch_dependents();
  extern __shared__ uint8_t gnt_raw[];
  __shared__ uint64_t bar[2];
  __shared__ float red_s[2 * GNF_THREADS], red_q[2 * GNF_THREADS];
  __shared__ float tot_s[2 * GNF_MAX_CH / 8], tot_q[2 * GNF_MAX_CH / 8];
  __shared__ float2 part[8];
  __shared__ float s_mean[8], s_rstd[8];
  const int tid = threadIdx.x;
  const int R = gridDim.x, rank = blockIdx.x;
  const int row_begin = rank * rows_cta;
  const int nrows = min(HW, row_begin + rows_cta) - row_begin;
  const uint32_t slab0 = (smem_u32(gnt_raw) + 127u) & ~127u;
  const uint32_t slab_stride = (static_cast<uint32_t>(slab_bytes) + 127u) & ~127u;
  if (tid == 0) {
    mbar_init(&bar[0], 1);
    mbar_init(&bar[1], 1);
    fence_barrier_init();
  }
  __syncthreads();
  pdl_wait();
  const int VPR = CH >> 3;
  const int rows_par = GNF_THREADS / VPR;
  const int vc = tid % VPR, rsub = tid / VPR;
  const bool active = rsub < rows_par;
  const int ngc = CH / cpg;
  const int mine = nrows - rsub;
  const int nvalid = (active && mine > 0) ? (mine + rows_par - 1) / rows_par : 0;
  const uint32_t sstep = static_cast<uint32_t>(rows_par * VPR * 16);

  auto issue = [&](int item, int buf) {  // one elected thread of warp 0
    const int chunk = item % chunks, img = item / chunks;
    mbar_arrive_expect_tx(&bar[buf], static_cast<uint32_t>(nbox * boxr * CH * 2));
    for (int b = 0; b < nbox; ++b) {
      asm volatile(
          "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
          ::"r"(slab0 + static_cast<uint32_t>(buf) * slab_stride + static_cast<uint32_t>(b * boxr * CH * 2)),
          "l"(reinterpret_cast<uint64_t>(&tmx)), "r"(smem_u32(&bar[buf])), "r"(chunk * CH), "r"(row_begin + b * boxr),
          "r"(img)
          : "memory");
    }
  };
  const int first = blockIdx.y, step = gridDim.y;
  if (tid < 32 && first < items) {
    if (elect_one()) issue(first, 0);
  }
  int it = 0;
  for (int item = first; item < items; item += step, ++it) {
    const int buf = it & 1;
    const int chunk = item % chunks, img = item / chunks;
    // prefetch the next item's slab into the other buffer: every thread finished reading it before the __syncthreads
    // that ended the previous iteration
    if (tid < 32 && item + step < items) {
      if (elect_one()) issue(item + step, buf ^ 1);
    }
    const uint32_t my0 = slab0 + static_cast<uint32_t>(buf) * slab_stride + static_cast<uint32_t>((rsub * VPR + vc) * 16);
    const int c0 = chunk * CH + vc * 8;
    const int g_first = c0 / cpg;
    const int split = min(8, (g_first + 1) * cpg - c0);
    uint64_t rb2[4];
    if (rowbias != nullptr) {
      const float4* rp = reinterpret_cast<const float4*>(rowbias + (img / rb_div) * ldrb + c0);
      const float4 r0 = __ldg(rp), r1 = __ldg(rp + 1);
      rb2[0] = f2_pack(r0.x, r0.y); rb2[1] = f2_pack(r0.z, r0.w); rb2[2] = f2_pack(r1.x, r1.y); rb2[3] = f2_pack(r1.z, r1.w);
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j) rb2[j] = 0ull;
    }
    mbar_wait(&bar[buf], static_cast<uint32_t>(it >> 1) & 1u);
    {
      uint64_t a2[4], b2[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) a2[j] = b2[j] = 0ull;
#pragma unroll 4
      for (int i = 0; i < nvalid; ++i) {
        uint32_t w[4];
        asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];"
                     : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]) : "r"(my0 + i * sstep));
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const uint64_t xr = f2_add(bf16x2_to_f2(w[j]), rb2[j]);
          a2[j] = f2_add(a2[j], xr);
          b2[j] = f2_fma(xr, xr, b2[j]);
        }
      }
      float a[8], b[8];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        f2_unpack(a2[j], a[2 * j], a[2 * j + 1]);
        f2_unpack(b2[j], b[2 * j], b[2 * j + 1]);
      }
      float s_lo = 0.f, q_lo = 0.f, s_hi = 0.f, q_hi = 0.f;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        if (j < split) { s_lo += a[j]; q_lo += b[j]; } else { s_hi += a[j]; q_hi += b[j]; }
      }
      if (active) {
        red_s[(vc * 2) * rows_par + rsub] = s_lo;
        red_q[(vc * 2) * rows_par + rsub] = q_lo;
        red_s[(vc * 2 + 1) * rows_par + rsub] = s_hi;
        red_q[(vc * 2 + 1) * rows_par + rsub] = q_hi;
      }
    }
    __syncthreads();
    {
      const int warp = tid >> 5, lane = tid & 31;
      for (int s = warp; s < 2 * VPR; s += GNF_THREADS / 32) {
        float ss = 0.f, qq = 0.f;
        for (int r = lane; r < rows_par; r += 32) {
          ss += red_s[s * rows_par + r];
          qq += red_q[s * rows_par + r];
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          ss += __shfl_xor_sync(0xffffffffu, ss, o);
          qq += __shfl_xor_sync(0xffffffffu, qq, o);
        }
        if (lane == 0) { tot_s[s] = ss; tot_q[s] = qq; }
      }
    }
    __syncthreads();
    if (tid < ngc) {
      const int g = chunk * ngc + tid;
      float ss = 0.f, qq = 0.f;
      for (int s = 0; s < 2 * VPR; ++s) {
        const int sc0 = chunk * CH + (s >> 1) * 8;
        if (sc0 / cpg + (s & 1) == g) { ss += tot_s[s]; qq += tot_q[s]; }
      }
      part[tid] = make_float2(ss, qq);
    }
    if (R > 1) { cluster_arrive_release(); cluster_wait_acquire(); } else { __syncthreads(); }
    if (tid < ngc) {
      float ss = 0.f, qq = 0.f;
      if (R > 1) {
        const uint32_t mine_addr = smem_u32(&part[tid]);
        float2 p[GNF_MAX_CLUSTER];
#pragma unroll
        for (int r = 0; r < GNF_MAX_CLUSTER; ++r)
          p[r] = r < R ? ld_dsmem_f2(mapa_shared(mine_addr, static_cast<uint32_t>(r))) : make_float2(0.f, 0.f);
#pragma unroll
        for (int r = 0; r < GNF_MAX_CLUSTER; ++r) {
          if (r < R) { ss += p[r].x; qq += p[r].y; }
        }
      } else {
        ss = part[tid].x; qq = part[tid].y;
      }
      const float inv_n = 1.0f / (static_cast<float>(HW) * cpg);
      const float mean = ss * inv_n;
      const float var = fmaxf(qq * inv_n - mean * mean, 0.f);
      s_mean[tid] = mean;
      s_rstd[tid] = rsqrtf(var + eps);
    }
    __syncthreads();
    if (R > 1) cluster_arrive_release();  // my remote reads of every CTA's `part` are done

    if (nvalid > 0) {
      float sa[8], sb[8], rbf[8];
#pragma unroll
      for (int j = 0; j < 8; j += 4) {
        const float4 g4 = __ldg(reinterpret_cast<const float4*>(gamma + c0 + j));
        const float4 b4 = __ldg(reinterpret_cast<const float4*>(beta + c0 + j));
        sa[j] = g4.x; sa[j + 1] = g4.y; sa[j + 2] = g4.z; sa[j + 3] = g4.w;
        sb[j] = b4.x; sb[j + 1] = b4.y; sb[j + 2] = b4.z; sb[j + 3] = b4.w;
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) f2_unpack(rb2[j], rbf[2 * j], rbf[2 * j + 1]);
      const int gl = g_first - chunk * ngc;
      const float m0 = s_mean[gl], r0 = s_rstd[gl];
      const float m1 = split < 8 ? s_mean[gl + 1] : 0.f, r1 = split < 8 ? s_rstd[gl + 1] : 0.f;
      uint64_t sa2[4], sb2[4];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float a = (j < split ? r0 : r1) * sa[j];
        sb[j] = fmaf(rbf[j] - (j < split ? m0 : m1), a, sb[j]);
        sa[j] = a;
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        sa2[j] = f2_pack(sa[2 * j], sa[2 * j + 1]);
        sb2[j] = f2_pack(sb[2 * j], sb[2 * j + 1]);
      }
      uint4* op = reinterpret_cast<uint4*>(out + (static_cast<long long>(img) * HW + row_begin + rsub) * ldo + c0);
      const long long ostep = static_cast<long long>(rows_par) * ldo / 8;
#pragma unroll 4
      for (int i = 0; i < nvalid; ++i) {
        uint32_t w[4], o[4];
        asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];"
                     : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]) : "r"(my0 + i * sstep));
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          uint64_t y2 = f2_fma(bf16x2_to_f2(w[j]), sa2[j], sb2[j]);
          if (silu) y2 = silu_f2(y2);
          o[j] = f2_to_bf16x2(y2);
        }
        op[i * ostep] = make_uint4(o[0], o[1], o[2], o[3]);
      }
    }
    // every CTA of the cluster has read my `part` (second cluster barrier) and every thread of this CTA is done with
    // slab `buf`, s_mean / s_rstd and the reduction scratch before the next iteration overwrites them
    if (R > 1) cluster_wait_acquire();
    __syncthreads();
  }
}

// Shape plan of the single-pass kernel: returns 0 when the shape does not fit (the three-kernel form takes it).
struct GnFusedPlan {
  int ch, cluster, rows_cta, vmax;
};
static bool gn_fused_plan(int HW, int C, int G, GnFusedPlan* p) {
  const int cpg = C / G;
  if (cpg < 8) return false;  // a 16-byte vector must not span more than two groups
  int g8 = 8, a = cpg;
  while (a) { const int t = g8 % a; g8 = a; a = t; }  // gcd(cpg, 8)
  const int ch = cpg * (8 / g8);
  if (ch > GNF_MAX_CH || C % ch != 0 || ch / cpg > 8) return false;
  const int rows_par = GNF_THREADS / (ch / 8);
  for (int r = 1; r <= GNF_MAX_CLUSTER; r *= 2) {
    const int rows_cta = ceil_div(HW, r);
    const int need = ceil_div(rows_cta, rows_par);
    if (need <= 8) {
      p->ch = ch; p->cluster = r; p->rows_cta = rows_cta; p->vmax = need <= 2 ? 2 : need <= 4 ? 4 : 8;
      return true;
    }
  }
  return false;
}

struct GnTmaPlan {
  int ch, cluster, rows_cta, boxr, nbox, slab;
};
static bool gn_tma_plan(int HW, int C, int G, int max_ch, GnTmaPlan* p) {
  const int cpg = C / G;
  if (cpg < 8) return false;
  int g8 = 8, a = cpg;
  while (a) { const int t = g8 % a; g8 = a; a = t; }
  const int ch = cpg * (8 / g8);
  if (ch > max_ch || ch > GNF_MAX_CH || C % ch != 0 || ch / cpg > 8) return false;
  for (int r = 1; r <= GNF_MAX_CLUSTER; r *= 2) {
    const int rows_cta = ceil_div(HW, r);
    const int nbox = ceil_div(rows_cta, 256);
    const int boxr = (ceil_div(rows_cta, nbox) + 7) / 8 * 8;  // box starts stay 128-byte aligned in the slab
    const int slab = nbox * boxr * ch * 2;
    if (boxr <= 256 && slab <= GNT_MAX_SLAB) {
      if (slab < GNT_MIN_SLAB) return false;
      p->ch = ch; p->cluster = r; p->rows_cta = rows_cta; p->boxr = boxr; p->nbox = nbox; p->slab = slab;
      return true;
    }
  }
  return false;
}

static int launch_gn_tma(const GnTmaPlan& pl, const void* x, long long ldx, const float* gamma, const float* beta,
                         float eps, void* out, long long ldo, int images, int HW, int C, int groups, int silu,
                         const float* rowbias, long long ldrb, int rb_div, cudaStream_t stream) {
  CUtensorMap tmx;
  const uint64_t dims[3] = {static_cast<uint64_t>(C), static_cast<uint64_t>(HW), static_cast<uint64_t>(images)};
  const uint64_t strides[2] = {static_cast<uint64_t>(ldx) * 2, static_cast<uint64_t>(HW) * static_cast<uint64_t>(ldx) * 2};
  const uint32_t box[3] = {static_cast<uint32_t>(pl.ch), static_cast<uint32_t>(pl.boxr), 1u};
  const int rc = make_tmap_bf16(&tmx, x, 3, dims, strides, box, false);
  if (rc != FMC_OK) return rc;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(pl.cluster, C / pl.ch, images);
  cfg.blockDim = dim3(GNF_THREADS);
  cfg.dynamicSmemBytes = static_cast<size_t>(pl.slab) + 128;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  int na = 0;
  if (pl.cluster > 1) {
    attr[na].id = cudaLaunchAttributeClusterDimension;
    attr[na].val.clusterDim.x = pl.cluster;
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
  FMC_CUDA_OK(cudaLaunchKernelEx(&cfg, groupnorm_tma_kernel, tmx, gamma, beta, eps, static_cast<__nv_bfloat16*>(out), ldo,
                                 HW, C / groups, pl.ch, pl.rows_cta, pl.boxr, pl.nbox, silu, rowbias, ldrb, rb_div));
  return check_launch("groupnorm_tma_kernel");
}

static int launch_gn_tma_persist(const GnTmaPlan& pl, const void* x, long long ldx, const float* gamma, const float* beta,
                                 float eps, void* out, long long ldo, int images, int HW, int C, int groups, int silu,
                                 const float* rowbias, long long ldrb, int rb_div, cudaStream_t stream) {
  CUtensorMap tmx;
  const uint64_t dims[3] = {static_cast<uint64_t>(C), static_cast<uint64_t>(HW), static_cast<uint64_t>(images)};
  const uint64_t strides[2] = {static_cast<uint64_t>(ldx) * 2, static_cast<uint64_t>(HW) * static_cast<uint64_t>(ldx) * 2};
  const uint32_t box[3] = {static_cast<uint32_t>(pl.ch), static_cast<uint32_t>(pl.boxr), 1u};
  const int rc = make_tmap_bf16(&tmx, x, 3, dims, strides, box, false);
  if (rc != FMC_OK) return rc;
  const int chunks = C / pl.ch;
  const int items = chunks * images;
  const int slab_stride = (pl.slab + 127) & ~127;
  const size_t smem = static_cast<size_t>(2) * slab_stride + 128;
  static unsigned long long attr_devs = 0;  // per device: the attribute belongs to the (device, function) pair
  if (first_use_on_this_device(&attr_devs)) {
    FMC_CUDA_OK(cudaFuncSetAttribute(groupnorm_tma_persist_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * (GNT_MAX_SLAB + 128) + 128));
  }
  cudaLaunchConfig_t cfg = {};
  cfg.blockDim = dim3(GNF_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  int na = 0;
  attr[na].id = cudaLaunchAttributeClusterDimension;
  attr[na].val.clusterDim.x = pl.cluster;
  attr[na].val.clusterDim.y = 1;
  attr[na].val.clusterDim.z = 1;
  ++na;
  // resident clusters for this (cluster size, shared memory) configuration: cached per configuration
  static int cached_key = -1, cached_clusters = 0;
  const int key = (current_device_ordinal() & 63) * 100000000 + pl.cluster * 1000000 + static_cast<int>(smem / 128);
  if (key != cached_key) {
    cfg.gridDim = dim3(pl.cluster, 1, 1);
    cfg.attrs = attr;
    cfg.numAttrs = na;
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, groupnorm_tma_persist_kernel, &cfg) != cudaSuccess || n <= 0) {
      cudaGetLastError();
      n = device_sm_count() * 2 / pl.cluster;
    }
    cached_key = key;
    cached_clusters = n;
  }
  const int nclusters = items < cached_clusters ? items : cached_clusters;
  cfg.gridDim = dim3(pl.cluster, nclusters, 1);
  if (pdl_enabled()) {
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  cfg.attrs = attr;
  cfg.numAttrs = na;
  FMC_CUDA_OK(cudaLaunchKernelEx(&cfg, groupnorm_tma_persist_kernel, tmx, gamma, beta, eps, static_cast<__nv_bfloat16*>(out),
                                 ldo, HW, C / groups, pl.ch, pl.rows_cta, pl.boxr, pl.nbox, silu, rowbias, ldrb, rb_div,
                                 chunks, items, pl.slab));
  return check_launch("groupnorm_tma_persist_kernel");
}

static int gn_fused_mode() {
  const char* e = getenv("FMC_GN_FUSED");  // read per call (A/B inside one process); a captured graph keeps its choice
  return e != nullptr ? (e[0] - '0') : GN_FUSED_DEFAULT;
}
// mode 0: never; 1: 40- and 80-channel chunks only (the 120-channel chunks of C = 960 / 1920 measured slower than the
// three-kernel form: 255 of 256 threads on 240-byte row pieces); 2: every shape that fits; 3 / 4: like 1 / 2 with the
// TMA-staged kernel wherever its slab is 8 - 32 KB (the register kernel keeps the small shapes)
static bool gn_fused_wanted(int HW, int C, int groups, GnFusedPlan* pl) {
  const int mode = gn_fused_mode();
  if (mode <= 0 || !gn_fused_plan(HW, C, groups, pl)) return false;
  return mode == 2 || mode == 4 || pl->ch <= 80;
}

template <int VMAX>
static cudaError_t launch_gn_fused(const GnFusedPlan& pl, const void* x, long long ldx, const float* gamma,
                                   const float* beta, float eps, void* out, long long ldo, int images, int HW, int C,
                                   int groups, int silu, const float* rowbias, long long ldrb, int rb_div,
                                   cudaStream_t stream) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(pl.cluster, C / pl.ch, images);
  cfg.blockDim = dim3(GNF_THREADS);
  cfg.dynamicSmemBytes = 0;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  int na = 0;
  if (pl.cluster > 1) {
    attr[na].id = cudaLaunchAttributeClusterDimension;
    attr[na].val.clusterDim.x = pl.cluster;
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
  return cudaLaunchKernelEx(&cfg, groupnorm_fused_kernel<VMAX>, static_cast<const __nv_bfloat16*>(x), ldx, gamma, beta,
                            eps, static_cast<__nv_bfloat16*>(out), ldo, HW, C / groups, pl.ch, pl.rows_cta, silu,
                            rowbias, ldrb, rb_div);
}

// Row statistics only: stats[row] = (mean, rstd) over C channels.  The LayerNorm itself is then applied inside the
// consuming GEMM (fmc_gemm_ln_bf16): one read pass instead of a read + write pass and a second read by the GEMM.
template <int NV>
__global__ void __launch_bounds__(LN_WARPS * 32)
rowstats_kernel(const __nv_bfloat16* __restrict__ x, long long ldx, float2* __restrict__ stats, long long rows, int C,
                float eps) {
  pdl_launch_dependents();
  pdl_wait();
  constexpr int R = 2;
  const int lane = threadIdx.x & 31;
  const long long row0 = (blockIdx.x * static_cast<long long>(LN_WARPS) + (threadIdx.x >> 5)) * R;
  if (row0 >= rows) return;
  const int nvec = C >> 3;
  const float inv_c = 1.0f / static_cast<float>(C);
  uint4 raw[R][NV];
#pragma unroll
  for (int r = 0; r < R; ++r) {
    const long long row = row0 + r < rows ? row0 + r : rows - 1;
    const uint4* xr = reinterpret_cast<const uint4*>(x + row * ldx);
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int vi = lane + i * 32;
      raw[r][i] = vi < nvec ? __ldg(xr + vi) : make_uint4(0u, 0u, 0u, 0u);
    }
  }
#pragma unroll
  for (int r = 0; r < R; ++r) {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const uint32_t w[4] = {raw[r][i].x, raw[r][i].y, raw[r][i].z, raw[r][i].w};
#pragma unroll
      for (int j = 0; j < 4; ++j) s += bf16_lo(w[j]) + bf16_hi(w[j]);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float mean = s * inv_c;
    float ss = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      if (lane + i * 32 < nvec) {
        const uint32_t w[4] = {raw[r][i].x, raw[r][i].y, raw[r][i].z, raw[r][i].w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float a = bf16_lo(w[j]) - mean, b = bf16_hi(w[j]) - mean;
          ss = fmaf(a, a, fmaf(b, b, ss));
        }
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
    if (lane == 0 && row0 + r < rows) stats[row0 + r] = make_float2(mean, rsqrtf(ss * inv_c + eps));
  }
}

template <int NV, int R>
static int launch_layernorm(const void* x, long long ldx, const float* gamma, const float* beta, float eps, void* out,
                            long long ldo, const float* pe, int F, int HW, const void* add, long long ldadd, void* out2,
                            long long ldo2, long long rows, int C, cudaStream_t stream) {
  const long long per_block = static_cast<long long>(LN_WARPS) * R;
  const unsigned grid = static_cast<unsigned>((rows + per_block - 1) / per_block);
  if (out2 != nullptr) {
    launch_k(layernorm_kernel<NV, R, true>, dim3(grid), dim3(LN_WARPS * 32), 0, stream, 
        static_cast<const __nv_bfloat16*>(x), ldx, gamma, beta, eps, static_cast<__nv_bfloat16*>(out), ldo, pe,
        F > 0 ? F : 1, HW > 0 ? HW : 1, static_cast<const __nv_bfloat16*>(add), ldadd,
        static_cast<__nv_bfloat16*>(out2), ldo2, rows, C);
  } else {
    launch_k(layernorm_kernel<NV, R, false>, dim3(grid), dim3(LN_WARPS * 32), 0, stream, 
        static_cast<const __nv_bfloat16*>(x), ldx, gamma, beta, eps, static_cast<__nv_bfloat16*>(out), ldo, pe,
        F > 0 ? F : 1, HW > 0 ? HW : 1, nullptr, 0, nullptr, 0, rows, C);
  }
  return check_launch("layernorm_kernel");
}

}  // namespace fmc

using namespace fmc;

extern "C" int fmc_layernorm_bf16(const void* x, long long ldx, const float* gamma, const float* beta, float eps,
                                  void* out, long long ldo, const float* pe, int F, int HW, const void* add,
                                  long long ldadd, void* out2, long long ldo2, long long rows, int C, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  FMC_REQUIRE(x && gamma && beta && out, FMC_ERR_ARG, "fmc_layernorm_bf16: null operand");
  FMC_REQUIRE(C % 8 == 0 && C <= 32 * 8 * LN_MAX_VEC, FMC_ERR_SHAPE, "fmc_layernorm_bf16: C=%d must be a multiple of 8 and <= %d",
              C, 32 * 8 * LN_MAX_VEC);
  FMC_REQUIRE(ldx % 8 == 0 && ldo % 8 == 0, FMC_ERR_SHAPE, "fmc_layernorm_bf16: row strides must be multiples of 8");
  FMC_REQUIRE((out2 == nullptr) == (add == nullptr), FMC_ERR_ARG, "fmc_layernorm_bf16: add and out2 go together");
  FMC_REQUIRE(pe == nullptr || (F > 0 && HW > 0), FMC_ERR_ARG, "fmc_layernorm_bf16: pe needs F and HW");
  if (rows == 0) return FMC_OK;
  // row passes in flight per warp: 2, except the two-output form (x + pose): it holds twice the vectors and lands at
  // 169 - 232 registers with 2 (one 8-warp block per SM, 2.9 TB/s); with 1 it needs 98 (two blocks, 3.95 TB/s).
  // FMC_LN_ADD_R=2 restores the old choice (A/B switch)
  int add_r = LN_ADD_R_DEFAULT;
  if (const char* e = getenv("FMC_LN_ADD_R")) add_r = atoi(e);
  if (add != nullptr && add_r == 1) {
    if (C == 320) return launch_layernorm40<8, 1>(x, ldx, gamma, beta, eps, out, ldo, pe, F, HW, add, ldadd, out2, ldo2, rows, stream);
    if (C == 640) return launch_layernorm40<16, 1>(x, ldx, gamma, beta, eps, out, ldo, pe, F, HW, add, ldadd, out2, ldo2, rows, stream);
    if (C == 1280) return launch_layernorm40<32, 1>(x, ldx, gamma, beta, eps, out, ldo, pe, F, HW, add, ldadd, out2, ldo2, rows, stream);
  }
  if (C == 320) return launch_layernorm40<8, 2>(x, ldx, gamma, beta, eps, out, ldo, pe, F, HW, add, ldadd, out2, ldo2, rows, stream);
  if (C == 640) return launch_layernorm40<16, 2>(x, ldx, gamma, beta, eps, out, ldo, pe, F, HW, add, ldadd, out2, ldo2, rows, stream);
  if (C == 1280) return launch_layernorm40<32, 2>(x, ldx, gamma, beta, eps, out, ldo, pe, F, HW, add, ldadd, out2, ldo2, rows, stream);
  const int nv = (C / 8 + 31) / 32;
#define FMC_LN_ARGS x, ldx, gamma, beta, eps, out, ldo, pe, F, HW, add, ldadd, out2, ldo2, rows, C, stream
  switch (nv) {
    case 1: return launch_layernorm<1, 4>(FMC_LN_ARGS);
    case 2: return launch_layernorm<2, 4>(FMC_LN_ARGS);
    case 3: return launch_layernorm<3, 2>(FMC_LN_ARGS);
    case 4: return launch_layernorm<4, 2>(FMC_LN_ARGS);
    default: return launch_layernorm<5, 2>(FMC_LN_ARGS);
  }
#undef FMC_LN_ARGS
}

extern "C" int fmc_groupnorm_bf16(const void* x, long long ldx, const float* gamma, const float* beta, float eps,
                                  void* out, long long ldo, float* stats_ws, int images, int HW, int C, int groups,
                                  int silu, const float* rowbias, long long ldrb, int rowbias_div, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  FMC_REQUIRE(x && gamma && beta && out && stats_ws, FMC_ERR_ARG, "fmc_groupnorm_bf16: null operand");
  FMC_REQUIRE(C % 8 == 0 && C % groups == 0 && groups <= 64 && C / 8 <= 1024, FMC_ERR_SHAPE,
              "fmc_groupnorm_bf16: unsupported C=%d groups=%d", C, groups);
  FMC_REQUIRE(ldx % 8 == 0 && ldo % 8 == 0, FMC_ERR_SHAPE, "fmc_groupnorm_bf16: row strides must be multiples of 8");
  if (images == 0 || HW == 0) return FMC_OK;
  const int rb_div0 = rowbias_div > 0 ? rowbias_div : 1;
  GnFusedPlan pl;
  GnTmaPlan tp;
  if (images <= 65535 && gn_fused_mode() >= 3 && (reinterpret_cast<uintptr_t>(x) & 15) == 0 &&
      gn_tma_plan(HW, C, groups, gn_fused_mode() >= 4 ? GNF_MAX_CH : 80, &tp)) {
    // mode 5: the persistent, double-buffered form (worth it only when a resident cluster gets several items)
    if (gn_fused_mode() >= 5 && static_cast<long long>(C / tp.ch) * images * tp.cluster > 2LL * device_sm_count() * 3)
      return launch_gn_tma_persist(tp, x, ldx, gamma, beta, eps, out, ldo, images, HW, C, groups, silu, rowbias, ldrb,
                                   rb_div0, stream);
    return launch_gn_tma(tp, x, ldx, gamma, beta, eps, out, ldo, images, HW, C, groups, silu, rowbias, ldrb, rb_div0, stream);
  }
  if (images <= 65535 && gn_fused_wanted(HW, C, groups, &pl)) {
    cudaError_t e;
    if (pl.vmax == 2) e = launch_gn_fused<2>(pl, x, ldx, gamma, beta, eps, out, ldo, images, HW, C, groups, silu, rowbias, ldrb, rb_div0, stream);
    else if (pl.vmax == 4) e = launch_gn_fused<4>(pl, x, ldx, gamma, beta, eps, out, ldo, images, HW, C, groups, silu, rowbias, ldrb, rb_div0, stream);
    else e = launch_gn_fused<8>(pl, x, ldx, gamma, beta, eps, out, ldo, images, HW, C, groups, silu, rowbias, ldrb, rb_div0, stream);
    FMC_CUDA_OK(e);
    return check_launch("groupnorm_fused_kernel");
  }
  // stats_ws: [images][chunks][groups][2] partial sums, then [images][C] (scale, shift) pairs
  const int chunks = ceil_div(HW, GN_ROWS);
  const int nvec = C / 8;
  const int threads = nvec * (nvec >= 512 ? 1 : 512 / nvec);  // whole rows per pass, <= 1024 threads
  const int rows_par = threads / nvec;
  const size_t smem = static_cast<size_t>(2) * rows_par * C * sizeof(float);
  static size_t smem_set_dev[64] = {0};
  size_t& smem_set = smem_set_dev[current_device_ordinal() & 63];
  if (smem > 48 * 1024 && smem > smem_set) {
    FMC_CUDA_OK(cudaFuncSetAttribute(groupnorm_partial_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     static_cast<int>(smem)));
    smem_set = smem;
  }
  const int rb_div = rowbias_div > 0 ? rowbias_div : 1;
  float2* ab = reinterpret_cast<float2*>(stats_ws + static_cast<size_t>(2) * groups * images * chunks);
  launch_k(groupnorm_partial_kernel, dim3(dim3(chunks, images)), dim3(threads), smem, stream, 
      static_cast<const __nv_bfloat16*>(x), ldx, stats_ws, HW, C, groups, chunks, rowbias, ldrb, rb_div);
  int rc = check_launch("groupnorm_partial_kernel");
  if (rc != FMC_OK) return rc;
  launch_k(groupnorm_finalize_kernel, dim3(images), dim3(256), 0, stream, stats_ws, gamma, beta, eps, ab, HW, C, groups, chunks, rowbias,
                                                        ldrb, rb_div);
  rc = check_launch("groupnorm_finalize_kernel");
  if (rc != FMC_OK) return rc;
  launch_k(groupnorm_apply_kernel, dim3(dim3(ceil_div(HW, rows_par * GN_UNROLL), images)), dim3(threads), 0, stream, 
      static_cast<const __nv_bfloat16*>(x), ldx, ab, static_cast<__nv_bfloat16*>(out), ldo, HW, C, silu);
  return check_launch("groupnorm_apply_kernel");
}

extern "C" int fmc_rowstats_bf16(const void* x, long long ldx, void* stats, long long rows, int C, float eps,
                                 void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  FMC_REQUIRE(x && stats, FMC_ERR_ARG, "fmc_rowstats_bf16: null operand");
  FMC_REQUIRE(C % 8 == 0 && C <= 32 * 8 * LN_MAX_VEC && ldx % 8 == 0, FMC_ERR_SHAPE,
              "fmc_rowstats_bf16: C=%d must be a multiple of 8 and <= %d, ldx a multiple of 8", C, 32 * 8 * LN_MAX_VEC);
  if (rows == 0) return FMC_OK;
  const unsigned grid = static_cast<unsigned>((rows + LN_WARPS * 2 - 1) / (LN_WARPS * 2));
  const __nv_bfloat16* xp = static_cast<const __nv_bfloat16*>(x);
  float2* sp = static_cast<float2*>(stats);
  const int nv = (C / 8 + 31) / 32;
  switch (nv) {
    case 1: FMC_CUDA_OK(launch_k(rowstats_kernel<1>, dim3(grid), dim3(LN_WARPS * 32), 0, stream, xp, ldx, sp, rows, C, eps)); break;
    case 2: FMC_CUDA_OK(launch_k(rowstats_kernel<2>, dim3(grid), dim3(LN_WARPS * 32), 0, stream, xp, ldx, sp, rows, C, eps)); break;
    case 3: FMC_CUDA_OK(launch_k(rowstats_kernel<3>, dim3(grid), dim3(LN_WARPS * 32), 0, stream, xp, ldx, sp, rows, C, eps)); break;
    case 4: FMC_CUDA_OK(launch_k(rowstats_kernel<4>, dim3(grid), dim3(LN_WARPS * 32), 0, stream, xp, ldx, sp, rows, C, eps)); break;
    default: FMC_CUDA_OK(launch_k(rowstats_kernel<5>, dim3(grid), dim3(LN_WARPS * 32), 0, stream, xp, ldx, sp, rows, C, eps)); break;
  }
  return check_launch("rowstats_kernel");
}

/* number of kernels one fmc_groupnorm_bf16 call of this shape launches (1: single-pass cluster kernel, 3: partial sums +
 * finalize + apply) -- for launch accounting only */
extern "C" int fmc_groupnorm_launches(int HW, int C, int groups) {
  GnFusedPlan pl;
  if (groups <= 0 || C % groups != 0) return 3;
  return gn_fused_wanted(HW, C, groups, &pl) ? 1 : 3;
}
