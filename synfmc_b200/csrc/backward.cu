// Backward kernels of the denoising hot path (train_cam_ctrl.py:586-665 `scaler.scale(loss).backward()`,
// train_cam_obj_ctrl.py:843-862): activation gradients through the frozen U-Net down to the CameraAdapter `qkv_merge`
// layers / the pose features (CMC) and to the injected object features (OMC), plus the pieces the parameter gradients of
// the trainable subset need.  bf16 tensors in the forward layout (channels-last rows), fp32 arithmetic inside.
// Linear layers need no kernel of their own: dX = dY W is fmc_gemm_bf16 against the transposed weight copy, dW = dY^T X
// is fmc_gemm_bf16 (fp32 output) on transposed operands (fmc_transpose_bf16), the bias gradient is fmc_colsum_f32.
// First version: correct and coalesced, not tuned -- the attention backward runs on CUDA cores.
#include <cuda_bf16.h>

#include "common.cuh"
#include "ptx.cuh"

namespace fmc {

__device__ __forceinline__ float bf2f(__nv_bfloat16 v) { return __bfloat162float(v); }
__device__ __forceinline__ void unpack8b(const uint4& u, float (&v)[8]) {
  const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    v[2 * j] = bf16_lo(w[j]);
    v[2 * j + 1] = bf16_hi(w[j]);
  }
}
__device__ __forceinline__ uint4 pack8b(const float (&v)[8]) {
  return make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]));
}
static inline unsigned bblocks(long long n) { return static_cast<unsigned>((n + 255) / 256); }

// ---------------------------------------------------------------------------------------------------------------
// out[c, r] = x[r, c]  (32 x 32 tiles through shared memory)
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
transpose_bf16_kernel(const __nv_bfloat16* __restrict__ x, long long ldx, __nv_bfloat16* __restrict__ out, long long ldo,
                      long long rows, int cols) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ __nv_bfloat16 tile[32][34];
  const long long r0 = static_cast<long long>(blockIdx.y) * 32;
  const int c0 = blockIdx.x * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int i = ty; i < 32; i += 8) {
    const long long r = r0 + i;
    const int c = c0 + tx;
    tile[i][tx] = (r < rows && c < cols) ? x[r * ldx + c] : __float2bfloat16(0.f);
  }
  __syncthreads();
  for (int i = ty; i < 32; i += 8) {
    const int c = c0 + i;
    const long long r = r0 + tx;
    if (c < cols && r < rows) out[static_cast<long long>(c) * ldo + r] = tile[tx][i];
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Column sums (bias gradients, LayerNorm parameter gradients): out[n] = sum_m x[m, n], fp32 accumulation, two stages
// in a fixed order (deterministic).  stage 1: grid (column blocks of 256, row chunks); stage 2 folds the chunks.
// ---------------------------------------------------------------------------------------------------------------
constexpr int COLSUM_ROWS = 256;
template <typename T>
__global__ void __launch_bounds__(256)
colsum_partial_kernel(const T* __restrict__ x, long long ldx, float* __restrict__ partial, long long rows, int cols) {
  pdl_launch_dependents();
  pdl_wait();
  const int c = blockIdx.x * 256 + threadIdx.x;
  if (c >= cols) return;
  const long long r0 = static_cast<long long>(blockIdx.y) * COLSUM_ROWS;
  const long long r1 = r0 + COLSUM_ROWS < rows ? r0 + COLSUM_ROWS : rows;
  float acc = 0.f;
  for (long long r = r0; r < r1; ++r) acc += static_cast<float>(x[r * ldx + c]);
  partial[static_cast<long long>(blockIdx.y) * cols + c] = acc;
}
__global__ void __launch_bounds__(256)
colsum_final_kernel(const float* __restrict__ partial, float* __restrict__ out, int chunks, int cols, int accumulate) {
  pdl_launch_dependents();
  pdl_wait();
  const int c = blockIdx.x * 256 + threadIdx.x;
  if (c >= cols) return;
  float acc = 0.f;
  for (int k = 0; k < chunks; ++k) acc += partial[static_cast<long long>(k) * cols + c];
  out[c] = accumulate ? out[c] + acc : acc;
}

// ---------------------------------------------------------------------------------------------------------------
// LayerNorm backward, one warp per row (the row in registers, C <= 1280):
//   xhat = (x - mean) rstd;  g = dy gamma;  dx = rstd (g - mean(g) - xhat mean(g xhat))
// optional parameter gradients: every block writes the sums of dy xhat and dy over ITS rows to
// pgrad[block][0 / 1][C]; the caller folds the blocks with fmc_colsum_f32.
// ---------------------------------------------------------------------------------------------------------------
constexpr int LNB_WARPS = 8;
constexpr int LNB_ROWS_PER_BLOCK = 64;
// NV: vectors of 8 channels per lane (C <= NV * 256); PARAMS: also accumulate d gamma / d beta (the accumulators cost 16 NV
// registers, so the frozen-LayerNorm form is a separate instantiation); R: rows a warp has in flight at once (the loads and
// the three shuffle reductions of R rows interleave -- a single row per warp leaves the memory system idle during its
// reduction chains: ncu showed 12 % of the DRAM peak at 22 % occupancy before).
template <int NV, bool PARAMS, int R>
__global__ void __launch_bounds__(LNB_WARPS * 32)
layernorm_bwd_kernel(const __nv_bfloat16* __restrict__ x, long long ldx, const __nv_bfloat16* __restrict__ dy, long long lddy,
                     const float* __restrict__ gamma, float eps, __nv_bfloat16* __restrict__ dx, long long lddx,
                     float* __restrict__ pgrad, long long rows, int C, int rows_per_block) {
  pdl_launch_dependents();
  pdl_wait();
  extern __shared__ float lnb_smem[];  // [LNB_WARPS][2][C] when PARAMS
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nvec = C >> 3;
  const float inv_c = 1.0f / static_cast<float>(C);
  float gm[NV][8];
  float dg[PARAMS ? NV : 1][8], db[PARAMS ? NV : 1][8];
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int vi = i * 32 + lane;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      gm[i][j] = vi < nvec ? __ldg(gamma + vi * 8 + j) : 0.f;
      if (PARAMS) dg[i][j] = db[i][j] = 0.f;
    }
  }
  const long long row0 = static_cast<long long>(blockIdx.x) * rows_per_block;
  for (int rr = warp * R; rr < rows_per_block; rr += LNB_WARPS * R) {
    if (row0 + rr >= rows) break;
    float xv[R][NV][8], gv[R][NV][8];
    float s[R], q[R], m1[R], m2[R], rstd[R];
    bool live[R];
#pragma unroll
    for (int k = 0; k < R; ++k) {
      const long long row = row0 + rr + k;
      live[k] = rr + k < rows_per_block && row < rows;
      s[k] = 0.f;
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const int vi = i * 32 + lane;
        if (live[k] && vi < nvec) {
          unpack8b(__ldg(reinterpret_cast<const uint4*>(x + row * ldx) + vi), xv[k][i]);
          unpack8b(__ldg(reinterpret_cast<const uint4*>(dy + row * lddy) + vi), gv[k][i]);
        } else {
#pragma unroll
          for (int j = 0; j < 8; ++j) xv[k][i][j] = gv[k][i][j] = 0.f;
        }
      }
    }
#pragma unroll
    for (int k = 0; k < R; ++k)
#pragma unroll
      for (int i = 0; i < NV; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) s[k] += xv[k][i][j];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
#pragma unroll
      for (int k = 0; k < R; ++k) s[k] += __shfl_xor_sync(0xffffffffu, s[k], o);
#pragma unroll
    for (int k = 0; k < R; ++k) {
      const float mean = s[k] * inv_c;
      q[k] = 0.f;
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const int vi = i * 32 + lane;
        if (vi < nvec) {
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            xv[k][i][j] -= mean;
            q[k] = fmaf(xv[k][i][j], xv[k][i][j], q[k]);
          }
        }
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
#pragma unroll
      for (int k = 0; k < R; ++k) q[k] += __shfl_xor_sync(0xffffffffu, q[k], o);
#pragma unroll
    for (int k = 0; k < R; ++k) {
      rstd[k] = 1.0f / sqrtf(q[k] * inv_c + eps);
      m1[k] = m2[k] = 0.f;
#pragma unroll
      for (int i = 0; i < NV; ++i) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          xv[k][i][j] *= rstd[k];                                 // xhat
          if (PARAMS) {
            dg[i][j] = fmaf(gv[k][i][j], xv[k][i][j], dg[i][j]);  // d gamma
            db[i][j] += gv[k][i][j];                               // d beta
          }
          gv[k][i][j] *= gm[i][j];                                // g = dy * gamma
          m1[k] += gv[k][i][j];
          m2[k] = fmaf(gv[k][i][j], xv[k][i][j], m2[k]);
        }
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
#pragma unroll
      for (int k = 0; k < R; ++k) {
        m1[k] += __shfl_xor_sync(0xffffffffu, m1[k], o);
        m2[k] += __shfl_xor_sync(0xffffffffu, m2[k], o);
      }
#pragma unroll
    for (int k = 0; k < R; ++k) {
      if (!live[k]) continue;
      const long long row = row0 + rr + k;
      const float a1 = m1[k] * inv_c, a2 = m2[k] * inv_c;
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const int vi = i * 32 + lane;
        if (vi < nvec) {
          float o8[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) o8[j] = rstd[k] * (gv[k][i][j] - a1 - xv[k][i][j] * a2);
          *(reinterpret_cast<uint4*>(dx + row * lddx) + vi) = pack8b(o8);
        }
      }
    }
  }
  if (PARAMS) {
    float* mine = lnb_smem + static_cast<size_t>(warp) * 2 * C;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int vi = i * 32 + lane;
      if (vi < nvec) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          mine[vi * 8 + j] = dg[i][j];
          mine[C + vi * 8 + j] = db[i][j];
        }
      }
    }
    __syncthreads();
    for (int c = threadIdx.x; c < 2 * C; c += LNB_WARPS * 32) {
      float acc = 0.f;
      for (int w = 0; w < LNB_WARPS; ++w) acc += lnb_smem[static_cast<size_t>(w) * 2 * C + c];
      pgrad[static_cast<long long>(blockIdx.x) * 2 * C + c] = acc;
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// GroupNorm (+ per-image channel bias before the norm, + SiLU after it) backward, frozen affine parameters:
//   u = x + rb;  xhat = (u - mu) rstd;  z = xhat gamma + beta;  y = silu(z) | z
//   dz = dy silu'(z);  g = dz gamma;  dx = rstd (g - mean_g(g) - xhat mean_g(g xhat))     (means over the group)
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float silu_grad(float z) {
  const float s = 1.0f / (1.0f + expf(-z));
  return s * (1.0f + z * (1.0f - s));
}
// Three passes over (image, row chunk) blocks, all with the same thread mapping -- a thread owns one 8-channel vector
// (fixed channels, so the per-channel constants sit in registers) and every RL-th row of the chunk, 16-byte loads:
//   pass 0: per-group chunk partials of sum u, sum u^2            -> part0[image, chunk, group, 2]
//   pass 1: mean / rstd from part0, then sum g, sum g xhat        -> part1
//   pass 2: both group means from the partials, then dx
// Partials are folded in a fixed order (deterministic, no atomics), the chunk totals in double.
constexpr int GNB_ROWS = 32;
constexpr int GNB_U = 4;  // rows a thread has in flight per trip
// d silu(z) / dz with the fast exponential / reciprocal (relative error ~1e-6, far below the bf16 output)
__device__ __forceinline__ float silu_grad_fast(float z) {
  const float s = __fdividef(1.0f, 1.0f + __expf(-z));
  return s * fmaf(z, 1.0f - s, 1.0f);
}
struct GnbParams {
  const __nv_bfloat16 *x, *dy;
  __nv_bfloat16* dx;
  long long ldx, lddy, lddx, ldrb;
  const float *gamma, *beta, *rowbias;
  float *part0, *part1;
  float eps;
  int images, HW, C, groups, silu, rb_div, chunks, chunk_rows;
};

template <int PASS>
__global__ void __launch_bounds__(256)
groupnorm_bwd_pass_kernel(GnbParams p) {
  pdl_launch_dependents();
  pdl_wait();
  extern __shared__ float gnb_part[];  // [RL][C][2] per-channel sums of this block (passes 0, 1)
  __shared__ double red[4][8][64];
  __shared__ float s_mean[64], s_rstd[64], s_a[64], s_b[64];
  const int tid = threadIdx.x;
  const int img = blockIdx.y, chunk = blockIdx.x;
  const int C = p.C, groups = p.groups, cg = C / groups, nvec = C >> 3;
  const int VT = nvec < 256 ? nvec : 256, RL = 256 / VT;
  const double inv_n = 1.0 / (static_cast<double>(p.HW) * cg);
  if (PASS >= 1) {
    const int nslice = 256 / groups > 8 ? 8 : 256 / groups;
    const int g = tid % groups, sl = tid / groups;
    if (sl < nslice) {
      double a0 = 0.0, a1 = 0.0, b0 = 0.0, b1 = 0.0;
      for (int ch = sl; ch < p.chunks; ch += nslice) {
        const long long o = ((static_cast<long long>(img) * p.chunks + ch) * groups + g) * 2;
        a0 += p.part0[o];
        a1 += p.part0[o + 1];
        if (PASS == 2) {
          b0 += p.part1[o];
          b1 += p.part1[o + 1];
        }
      }
      red[0][sl][g] = a0; red[1][sl][g] = a1; red[2][sl][g] = b0; red[3][sl][g] = b1;
    }
    __syncthreads();
    if (tid < groups) {
      double a0 = 0.0, a1 = 0.0, b0 = 0.0, b1 = 0.0;
      for (int q = 0; q < nslice; ++q) { a0 += red[0][q][tid]; a1 += red[1][q][tid]; b0 += red[2][q][tid]; b1 += red[3][q][tid]; }
      const double mean = a0 * inv_n;
      const double var = a1 * inv_n - mean * mean;
      s_mean[tid] = static_cast<float>(mean);
      s_rstd[tid] = 1.0f / sqrtf(static_cast<float>(var > 0.0 ? var : 0.0) + p.eps);
      s_a[tid] = static_cast<float>(b0 * inv_n);
      s_b[tid] = static_cast<float>(b1 * inv_n);
    }
    __syncthreads();
  }
  const int r0 = chunk * p.chunk_rows, r1 = min(p.HW, r0 + p.chunk_rows);
  const long long base = static_cast<long long>(img) * p.HW;
  const int vi0 = tid % VT, rl = tid / VT;
  if (rl < RL) {
    for (int vi = vi0; vi < nvec; vi += VT) {
      float gm[8], bt[8], rb[8], mu[8], rs[8], ga[8], gb[8], a[8], b[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int c = vi * 8 + j, g = c / cg;
        rb[j] = p.rowbias != nullptr ? __ldg(p.rowbias + static_cast<long long>(img / p.rb_div) * p.ldrb + c) : 0.f;
        a[j] = b[j] = 0.f;
        if (PASS >= 1) {
          gm[j] = __ldg(p.gamma + c); bt[j] = __ldg(p.beta + c);
          mu[j] = s_mean[g]; rs[j] = s_rstd[g];
        }
        if (PASS == 2) { ga[j] = s_a[g]; gb[j] = s_b[g]; }
      }
      // GNB_U rows per trip: their loads are issued together (memory-level parallelism; one row per trip left the kernel
      // latency-bound at 10 % of the DRAM peak), then the arithmetic of the rows runs on registers
      for (int rb0 = r0 + rl; rb0 < r1; rb0 += RL * GNB_U) {
        uint4 xq[GNB_U], dq[GNB_U];
#pragma unroll
        for (int k = 0; k < GNB_U; ++k) {
          const int r = rb0 + k * RL;
          if (r < r1) {
            xq[k] = __ldg(reinterpret_cast<const uint4*>(p.x + (base + r) * p.ldx) + vi);
            if (PASS >= 1) dq[k] = __ldg(reinterpret_cast<const uint4*>(p.dy + (base + r) * p.lddy) + vi);
          }
        }
#pragma unroll
        for (int k = 0; k < GNB_U; ++k) {
          const int r = rb0 + k * RL;
          if (r >= r1) break;
          float xv[8], dv[8], o8[8];
          unpack8b(xq[k], xv);
          if (PASS >= 1) unpack8b(dq[k], dv);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float u = xv[j] + rb[j];
            if (PASS == 0) {
              a[j] += u;
              b[j] = fmaf(u, u, b[j]);
            } else {
              const float xh = (u - mu[j]) * rs[j];
              float d = dv[j];
              if (p.silu) d *= silu_grad_fast(fmaf(xh, gm[j], bt[j]));
              const float g = d * gm[j];
              if (PASS == 1) {
                a[j] += g;
                b[j] = fmaf(g, xh, b[j]);
              } else {
                o8[j] = rs[j] * (g - ga[j] - xh * gb[j]);
              }
            }
          }
          if (PASS == 2) *(reinterpret_cast<uint4*>(p.dx + (base + r) * p.lddx) + vi) = pack8b(o8);
        }
      }
      if (PASS < 2) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          gnb_part[(static_cast<size_t>(rl) * C + vi * 8 + j) * 2] = a[j];
          gnb_part[(static_cast<size_t>(rl) * C + vi * 8 + j) * 2 + 1] = b[j];
        }
      }
    }
  }
  if (PASS < 2) {
    __syncthreads();
    float* out = PASS == 0 ? p.part0 : p.part1;
    for (int g = tid; g < groups; g += 256) {
      float sa = 0.f, sb = 0.f;
      for (int q = 0; q < RL; ++q)
        for (int c = g * cg; c < (g + 1) * cg; ++c) {
          sa += gnb_part[(static_cast<size_t>(q) * C + c) * 2];
          sb += gnb_part[(static_cast<size_t>(q) * C + c) * 2 + 1];
        }
      const long long o = ((static_cast<long long>(img) * p.chunks + chunk) * groups + g) * 2;
      out[o] = sa;
      out[o + 1] = sb;
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// GEGLU on the interleaved projection layout of the fused GEMM epilogue (blocks of 16 value | 16 gate columns):
//   forward  y[:, 16 b + j] = a gelu(g)                           (training forward: the projection is kept for backward)
//   backward dproj[value] = dy gelu(g);  dproj[gate] = dy a gelu'(g),  gelu'(g) = Phi(g) + g phi(g)   (exact erf GELU)
// ---------------------------------------------------------------------------------------------------------------
// Phi(g) = (1 + erf(g / sqrt 2)) / 2 with erf from Abramowitz & Stegun 7.1.28 (|err| <= 3e-7; the form the fused GEGLU
// epilogue of the inference GEMM uses, csrc/gemm.cu): 1 - erf(x) = (1 + a1 x + ... + a6 x^6)^-16 for x >= 0, the
// negative tail formed without cancellation.  One MUFU (rcp) per value instead of the ~40 instructions of erff.
__device__ __forceinline__ float normal_cdf_fast(float g) {
  const float x = fabsf(g) * 0.70710678118654752440f;
  float p = fmaf(x, 0.0000430638f, 0.0002765672f);
  p = fmaf(x, p, 0.0001520143f);
  p = fmaf(x, p, 0.0092705272f);
  p = fmaf(x, p, 0.0422820123f);
  p = fmaf(x, p, 0.0705230784f);
  p = fmaf(x, p, 1.0f);
  float r = rcp_fast(p);
  r *= r; r *= r; r *= r; r *= r;  // (1 / p)^16 = 1 - erf(x)
  return 0.5f * (g >= 0.f ? 2.0f - r : r);
}
__global__ void __launch_bounds__(256)
geglu_fwd_kernel(const __nv_bfloat16* __restrict__ proj, long long ldp, __nv_bfloat16* __restrict__ y, long long ldy,
                 long long rows, int H) {
  pdl_launch_dependents();
  pdl_wait();
  const int nvec = H >> 3;
  const long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (idx >= rows * nvec) return;
  const long long r = idx / nvec;
  const int vi = static_cast<int>(idx % nvec);
  const int blk = vi >> 1, half = vi & 1;  // 16-column block, which 8 of its 16 columns
  float a[8], g[8], o8[8];
  unpack8b(__ldg(reinterpret_cast<const uint4*>(proj + r * ldp + blk * 32 + half * 8)), a);
  unpack8b(__ldg(reinterpret_cast<const uint4*>(proj + r * ldp + blk * 32 + 16 + half * 8)), g);
#pragma unroll
  for (int j = 0; j < 8; ++j) o8[j] = a[j] * g[j] * normal_cdf_fast(g[j]);
  *(reinterpret_cast<uint4*>(y + r * ldy) + vi) = pack8b(o8);
}
__global__ void __launch_bounds__(256)
geglu_bwd_kernel(const __nv_bfloat16* __restrict__ proj, long long ldp, const __nv_bfloat16* __restrict__ dy, long long lddy,
                 __nv_bfloat16* __restrict__ dproj, long long lddp, long long rows, int H) {
  pdl_launch_dependents();
  pdl_wait();
  const int nvec = H >> 3;
  const long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (idx >= rows * nvec) return;
  const long long r = idx / nvec;
  const int vi = static_cast<int>(idx % nvec);
  const int blk = vi >> 1, half = vi & 1;
  float a[8], g[8], d[8], da[8], dg[8];
  unpack8b(__ldg(reinterpret_cast<const uint4*>(proj + r * ldp + blk * 32 + half * 8)), a);
  unpack8b(__ldg(reinterpret_cast<const uint4*>(proj + r * ldp + blk * 32 + 16 + half * 8)), g);
  unpack8b(__ldg(reinterpret_cast<const uint4*>(dy + r * lddy) + vi), d);
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float cdf = normal_cdf_fast(g[j]);  // gelu = g Phi(g);  gelu' = Phi(g) + g phi(g)
    da[j] = d[j] * g[j] * cdf;
    dg[j] = d[j] * a[j] * fmaf(g[j] * 0.39894228040143267794f, __expf(-0.5f * g[j] * g[j]), cdf);
  }
  *reinterpret_cast<uint4*>(dproj + r * lddp + blk * 32 + half * 8) = pack8b(da);
  *reinterpret_cast<uint4*>(dproj + r * lddp + blk * 32 + 16 + half * 8) = pack8b(dg);
}

// ---------------------------------------------------------------------------------------------------------------
// glue: ReLU, nearest upsample (integer factors), AvgPool2d(2)
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
relu_bwd_kernel(const __nv_bfloat16* __restrict__ y, const __nv_bfloat16* __restrict__ dy, __nv_bfloat16* __restrict__ dx,
                long long nvecs) {
  pdl_launch_dependents();
  pdl_wait();
  const long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (idx >= nvecs) return;
  float yv[8], dv[8];
  unpack8b(__ldg(reinterpret_cast<const uint4*>(y) + idx), yv);
  unpack8b(__ldg(reinterpret_cast<const uint4*>(dy) + idx), dv);
#pragma unroll
  for (int j = 0; j < 8; ++j) dv[j] = yv[j] > 0.f ? dv[j] : 0.f;
  *(reinterpret_cast<uint4*>(dx) + idx) = pack8b(dv);
}
// dx[n, y, x, :] = sum over the sy x sx output pixels that read input pixel (y, x)
__global__ void __launch_bounds__(256)
resize_nearest_bwd_kernel(const __nv_bfloat16* __restrict__ dy, __nv_bfloat16* __restrict__ dx, int N, int h, int w, int sy,
                          int sx, int C) {
  pdl_launch_dependents();
  pdl_wait();
  const int nvec = C >> 3;
  const long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  const long long total = static_cast<long long>(N) * h * w * nvec;
  if (idx >= total) return;
  const int vi = static_cast<int>(idx % nvec);
  long long t = idx / nvec;
  const int xx = static_cast<int>(t % w);
  t /= w;
  const int yy = static_cast<int>(t % h);
  const int n = static_cast<int>(t / h);
  const int oh = h * sy, ow = w * sx;
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  for (int a = 0; a < sy; ++a)
    for (int b = 0; b < sx; ++b) {
      float v[8];
      unpack8b(__ldg(reinterpret_cast<const uint4*>(
                         dy + ((static_cast<long long>(n) * oh + yy * sy + a) * ow + xx * sx + b) * C) + vi), v);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] += v[j];
    }
  *(reinterpret_cast<uint4*>(dx) + idx) = pack8b(acc);
}
__global__ void __launch_bounds__(256)
avgpool2_bwd_kernel(const __nv_bfloat16* __restrict__ dy, __nv_bfloat16* __restrict__ dx, int N, int h, int w, int C) {
  pdl_launch_dependents();
  pdl_wait();
  const int nvec = C >> 3;
  const long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  const long long total = static_cast<long long>(N) * h * w * nvec;
  if (idx >= total) return;
  const int vi = static_cast<int>(idx % nvec);
  long long t = idx / nvec;
  const int xx = static_cast<int>(t % w);
  t /= w;
  const int yy = static_cast<int>(t % h);
  const int n = static_cast<int>(t / h);
  const int oh = h >> 1, ow = w >> 1;
  float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if ((yy >> 1) < oh && (xx >> 1) < ow) {
    unpack8b(__ldg(reinterpret_cast<const uint4*>(dy + ((static_cast<long long>(n) * oh + (yy >> 1)) * ow + (xx >> 1)) * C) + vi), v);
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] *= 0.25f;
  }
  *(reinterpret_cast<uint4*>(dx) + idx) = pack8b(v);
}

// ---------------------------------------------------------------------------------------------------------------
// Attention backward on CUDA cores (flash style, probabilities recomputed; same row addressing as fmc_attention_f32:
// row of element t of sequence i = (i / inner) len inner + i % inner + t inner).  bf16 tensors, fp32 arithmetic.
//   P = softmax(Q K^T s);  dV = P^T dO;  dP = dO V^T;  dS = P (dP - D) s,  D_i = sum_c dO_ic O_ic;  dQ = dS K;  dK = dS^T Q
// kernel 1 (per 32 queries x head): row log-sum-exp L and D, then dQ; L and D go to lse[row, head] / dsum[row, head].
// kernel 2 (per 32 keys x head, self-attention only): dK, dV over all queries of the sequence, using L and D.
// Heads of Q / K start at col0 + h * head_stride (zero padding when head_dim = 40), heads of V / O / dO at col0 + h * D;
// dQ / dK are written in the Q / K layout, dV in the V layout (the [token, q|k|v] gradient buffer of the fused projection).
// ---------------------------------------------------------------------------------------------------------------
struct AttnBwdParams {
  const __nv_bfloat16 *Q, *K, *V, *O, *dO;
  __nv_bfloat16 *dQ, *dK, *dV;
  long long ldq, ldk, ldv, ldo, lddo, lddq, lddk, lddv;
  int q_col0, k_col0, v_col0, dq_col0, dk_col0, dv_col0, head_stride;
  float *lse, *dsum;  // [q_rows, heads]
  int images, heads, nq, nk, kv_div, kv_stride, inner;
  float scale;
};

template <int D>
__global__ void __launch_bounds__(128)
attention_bwd_dq_kernel(AttnBwdParams p) {
  pdl_launch_dependents();
  pdl_wait();
  constexpr int NCC = (D + 31) / 32;
  constexpr int KS = D + 1;
  extern __shared__ float smem_f[];
  float* Ks = smem_f;                 // [32][KS]
  float* Vs = Ks + 32 * KS;           // [32][KS]
  float* Qs = Vs + 32 * KS;           // [4][D][8]
  float* Gs = Qs + 4 * D * 8;         // [4][D][8]   dO
  float* Ps = Gs + 4 * D * 8;         // [4][32][8]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int img = blockIdx.z, head = blockIdx.y;
  const int q0 = blockIdx.x * 32 + warp * 8;
  const long long q_base = static_cast<long long>(img / p.inner) * p.nq * p.inner + (img % p.inner);
  const int g = img / p.kv_div;
  const long long kv_base = static_cast<long long>(g / p.inner) * p.kv_stride * p.inner + (g % p.inner);
  float* Qw = Qs + warp * D * 8;
  float* Gw = Gs + warp * D * 8;
  float* Pw = Ps + warp * 32 * 8;
  float dsum[8];
#pragma unroll
  for (int r = 0; r < 8; ++r) {
    const int t = min(q0 + r, p.nq - 1);
    const long long row = q_base + static_cast<long long>(t) * p.inner;
    const __nv_bfloat16* qp = p.Q + row * p.ldq + p.q_col0 + head * p.head_stride;
    const __nv_bfloat16* gp = p.dO + row * p.lddo + head * D;
    const __nv_bfloat16* op = p.O + row * p.ldo + head * D;
    float acc = 0.f;
    for (int c = lane; c < D; c += 32) {
      Qw[c * 8 + r] = bf2f(qp[c]) * p.scale;
      const float gv = bf2f(gp[c]);
      Gw[c * 8 + r] = gv;
      acc += gv * bf2f(op[c]);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    dsum[r] = acc;
  }
  float m[8], l[8], dq[8][NCC];
#pragma unroll
  for (int r = 0; r < 8; ++r) {
    m[r] = -INFINITY;
    l[r] = 0.f;
#pragma unroll
    for (int cc = 0; cc < NCC; ++cc) dq[r][cc] = 0.f;
  }
  auto load_tile = [&](int kt, bool with_v) {
    __syncthreads();
    for (int idx = threadIdx.x; idx < 32 * (D / 8); idx += 128) {
      const int j = idx / (D / 8), c8 = idx % (D / 8);
      float kv[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, vv[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      if (kt + j < p.nk) {
        const long long row = kv_base + static_cast<long long>(kt + j) * p.inner;
        unpack8b(__ldg(reinterpret_cast<const uint4*>(p.K + row * p.ldk + p.k_col0 + head * p.head_stride) + c8), kv);
        if (with_v) unpack8b(__ldg(reinterpret_cast<const uint4*>(p.V + row * p.ldv + p.v_col0 + head * D) + c8), vv);
      }
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        Ks[j * KS + c8 * 8 + e] = kv[e];
        Vs[j * KS + c8 * 8 + e] = vv[e];
      }
    }
    __syncthreads();
  };
  auto scores = [&](float (&s)[8]) {
#pragma unroll
    for (int r = 0; r < 8; ++r) s[r] = 0.f;
    const float* kr = Ks + lane * KS;
#pragma unroll 4
    for (int c = 0; c < D; ++c) {
      const float kv = kr[c];
      const float4 qa = *reinterpret_cast<const float4*>(Qw + c * 8);
      const float4 qb = *reinterpret_cast<const float4*>(Qw + c * 8 + 4);
      s[0] = fmaf(qa.x, kv, s[0]); s[1] = fmaf(qa.y, kv, s[1]); s[2] = fmaf(qa.z, kv, s[2]); s[3] = fmaf(qa.w, kv, s[3]);
      s[4] = fmaf(qb.x, kv, s[4]); s[5] = fmaf(qb.y, kv, s[5]); s[6] = fmaf(qb.z, kv, s[6]); s[7] = fmaf(qb.w, kv, s[7]);
    }
  };
  // ---- pass A: log-sum-exp of every query row
  for (int kt = 0; kt < p.nk; kt += 32) {
    load_tile(kt, false);
    float s[8];
    scores(s);
    const bool valid = kt + lane < p.nk;
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      const float sv = valid ? s[r] : -INFINITY;
      float mx = sv;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
      const float m_new = fmaxf(m[r], mx);
      float sum = valid ? expf(sv - m_new) : 0.f;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
      l[r] = l[r] * expf(m[r] - m_new) + sum;
      m[r] = m_new;
    }
  }
  float lse[8];
#pragma unroll
  for (int r = 0; r < 8; ++r) lse[r] = m[r] + logf(l[r]);
  // ---- pass B: dS and dQ
  for (int kt = 0; kt < p.nk; kt += 32) {
    load_tile(kt, true);
    float s[8];
    scores(s);
    float dp[8];
#pragma unroll
    for (int r = 0; r < 8; ++r) dp[r] = 0.f;
    const float* vr = Vs + lane * KS;
#pragma unroll 4
    for (int c = 0; c < D; ++c) {
      const float vv = vr[c];
      const float4 ga = *reinterpret_cast<const float4*>(Gw + c * 8);
      const float4 gb = *reinterpret_cast<const float4*>(Gw + c * 8 + 4);
      dp[0] = fmaf(ga.x, vv, dp[0]); dp[1] = fmaf(ga.y, vv, dp[1]); dp[2] = fmaf(ga.z, vv, dp[2]); dp[3] = fmaf(ga.w, vv, dp[3]);
      dp[4] = fmaf(gb.x, vv, dp[4]); dp[5] = fmaf(gb.y, vv, dp[5]); dp[6] = fmaf(gb.z, vv, dp[6]); dp[7] = fmaf(gb.w, vv, dp[7]);
    }
    const bool valid = kt + lane < p.nk;
    float ds[8];
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      const float pv = valid ? expf(s[r] - lse[r]) : 0.f;
      ds[r] = pv * (dp[r] - dsum[r]) * p.scale;
    }
    *reinterpret_cast<float4*>(Pw + lane * 8) = make_float4(ds[0], ds[1], ds[2], ds[3]);
    *reinterpret_cast<float4*>(Pw + lane * 8 + 4) = make_float4(ds[4], ds[5], ds[6], ds[7]);
    __syncwarp();
#pragma unroll 2
    for (int j = 0; j < 32; ++j) {
      const float4 pa = *reinterpret_cast<const float4*>(Pw + j * 8);
      const float4 pb = *reinterpret_cast<const float4*>(Pw + j * 8 + 4);
#pragma unroll
      for (int cc = 0; cc < NCC; ++cc) {
        const int c = cc * 32 + lane;
        const float kv = Ks[j * KS + (c < D ? c : 0)];
        dq[0][cc] = fmaf(pa.x, kv, dq[0][cc]); dq[1][cc] = fmaf(pa.y, kv, dq[1][cc]);
        dq[2][cc] = fmaf(pa.z, kv, dq[2][cc]); dq[3][cc] = fmaf(pa.w, kv, dq[3][cc]);
        dq[4][cc] = fmaf(pb.x, kv, dq[4][cc]); dq[5][cc] = fmaf(pb.y, kv, dq[5][cc]);
        dq[6][cc] = fmaf(pb.z, kv, dq[6][cc]); dq[7][cc] = fmaf(pb.w, kv, dq[7][cc]);
      }
    }
    __syncwarp();
  }
#pragma unroll
  for (int r = 0; r < 8; ++r) {
    const int t = q0 + r;
    if (t >= p.nq) continue;
    const long long row = q_base + static_cast<long long>(t) * p.inner;
    __nv_bfloat16* dst = p.dQ + row * p.lddq + p.dq_col0 + head * p.head_stride;
#pragma unroll
    for (int cc = 0; cc < NCC; ++cc) {
      const int c = cc * 32 + lane;
      if (c < D) dst[c] = __float2bfloat16(dq[r][cc]);
    }
    if (lane == 0) {
      p.lse[row * p.heads + head] = lse[r];
      p.dsum[row * p.heads + head] = dsum[r];
    }
  }
}

template <int D>
__global__ void __launch_bounds__(128)
attention_bwd_dkv_kernel(AttnBwdParams p) {
  pdl_launch_dependents();
  pdl_wait();
  constexpr int NCC = (D + 31) / 32;
  constexpr int KS = D + 1;
  extern __shared__ float smem_f[];
  float* Qs = smem_f;                 // [32][KS]  query tile (scaled)
  float* Gs = Qs + 32 * KS;           // [32][KS]  dO tile
  float* Kw_all = Gs + 32 * KS;       // [4][D][8] this warp's keys
  float* Vw_all = Kw_all + 4 * D * 8; // [4][D][8]
  float* Ps = Vw_all + 4 * D * 8;     // [4][32][8]
  float* Ls = Ps + 4 * 32 * 8;        // [32] lse, [32] dsum
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int img = blockIdx.z, head = blockIdx.y;
  const int k0 = blockIdx.x * 32 + warp * 8;
  const long long base = static_cast<long long>(img / p.inner) * p.nq * p.inner + (img % p.inner);  // self-attention: nq == nk
  float* Kw = Kw_all + warp * D * 8;
  float* Vw = Vw_all + warp * D * 8;
  float* Pw = Ps + warp * 32 * 8;
#pragma unroll
  for (int r = 0; r < 8; ++r) {
    const int t = min(k0 + r, p.nk - 1);
    const long long row = base + static_cast<long long>(t) * p.inner;
    const __nv_bfloat16* kp = p.K + row * p.ldk + p.k_col0 + head * p.head_stride;
    const __nv_bfloat16* vp = p.V + row * p.ldv + p.v_col0 + head * D;
    for (int c = lane; c < D; c += 32) {
      Kw[c * 8 + r] = bf2f(kp[c]);
      Vw[c * 8 + r] = bf2f(vp[c]);
    }
  }
  float dk[8][NCC], dv[8][NCC];
#pragma unroll
  for (int r = 0; r < 8; ++r)
#pragma unroll
    for (int cc = 0; cc < NCC; ++cc) dk[r][cc] = dv[r][cc] = 0.f;

  for (int qt = 0; qt < p.nq; qt += 32) {
    __syncthreads();
    for (int idx = threadIdx.x; idx < 32 * (D / 8); idx += 128) {
      const int j = idx / (D / 8), c8 = idx % (D / 8);
      float qv[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, gv[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      if (qt + j < p.nq) {
        const long long row = base + static_cast<long long>(qt + j) * p.inner;
        unpack8b(__ldg(reinterpret_cast<const uint4*>(p.Q + row * p.ldq + p.q_col0 + head * p.head_stride) + c8), qv);
        unpack8b(__ldg(reinterpret_cast<const uint4*>(p.dO + row * p.lddo + head * D) + c8), gv);
      }
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        Qs[j * KS + c8 * 8 + e] = qv[e] * p.scale;
        Gs[j * KS + c8 * 8 + e] = gv[e];
      }
    }
    if (threadIdx.x < 32) {
      const bool ok = qt + threadIdx.x < p.nq;
      const long long row = base + static_cast<long long>(ok ? qt + threadIdx.x : 0) * p.inner;
      Ls[threadIdx.x] = ok ? p.lse[row * p.heads + head] : INFINITY;  // exp(s - inf) = 0 for queries past the end
      Ls[32 + threadIdx.x] = ok ? p.dsum[row * p.heads + head] : 0.f;
    }
    __syncthreads();
    // lane = query of the tile, r = key of this warp
    float s[8], dp[8];
#pragma unroll
    for (int r = 0; r < 8; ++r) s[r] = dp[r] = 0.f;
    const float* qr = Qs + lane * KS;
    const float* gr = Gs + lane * KS;
#pragma unroll 4
    for (int c = 0; c < D; ++c) {
      const float qv = qr[c], gv = gr[c];
      const float4 ka = *reinterpret_cast<const float4*>(Kw + c * 8);
      const float4 kb = *reinterpret_cast<const float4*>(Kw + c * 8 + 4);
      const float4 va = *reinterpret_cast<const float4*>(Vw + c * 8);
      const float4 vb = *reinterpret_cast<const float4*>(Vw + c * 8 + 4);
      s[0] = fmaf(ka.x, qv, s[0]); s[1] = fmaf(ka.y, qv, s[1]); s[2] = fmaf(ka.z, qv, s[2]); s[3] = fmaf(ka.w, qv, s[3]);
      s[4] = fmaf(kb.x, qv, s[4]); s[5] = fmaf(kb.y, qv, s[5]); s[6] = fmaf(kb.z, qv, s[6]); s[7] = fmaf(kb.w, qv, s[7]);
      dp[0] = fmaf(va.x, gv, dp[0]); dp[1] = fmaf(va.y, gv, dp[1]); dp[2] = fmaf(va.z, gv, dp[2]); dp[3] = fmaf(va.w, gv, dp[3]);
      dp[4] = fmaf(vb.x, gv, dp[4]); dp[5] = fmaf(vb.y, gv, dp[5]); dp[6] = fmaf(vb.z, gv, dp[6]); dp[7] = fmaf(vb.w, gv, dp[7]);
    }
    const float lq = Ls[lane], dq_ = Ls[32 + lane];
    float pr[8], ds[8];
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      pr[r] = expf(s[r] - lq);
      ds[r] = pr[r] * (dp[r] - dq_) * p.scale;
    }
    // dV += P^T dO
    *reinterpret_cast<float4*>(Pw + lane * 8) = make_float4(pr[0], pr[1], pr[2], pr[3]);
    *reinterpret_cast<float4*>(Pw + lane * 8 + 4) = make_float4(pr[4], pr[5], pr[6], pr[7]);
    __syncwarp();
#pragma unroll 2
    for (int j = 0; j < 32; ++j) {
      const float4 pa = *reinterpret_cast<const float4*>(Pw + j * 8);
      const float4 pb = *reinterpret_cast<const float4*>(Pw + j * 8 + 4);
#pragma unroll
      for (int cc = 0; cc < NCC; ++cc) {
        const int c = cc * 32 + lane;
        const float gv = Gs[j * KS + (c < D ? c : 0)];
        dv[0][cc] = fmaf(pa.x, gv, dv[0][cc]); dv[1][cc] = fmaf(pa.y, gv, dv[1][cc]);
        dv[2][cc] = fmaf(pa.z, gv, dv[2][cc]); dv[3][cc] = fmaf(pa.w, gv, dv[3][cc]);
        dv[4][cc] = fmaf(pb.x, gv, dv[4][cc]); dv[5][cc] = fmaf(pb.y, gv, dv[5][cc]);
        dv[6][cc] = fmaf(pb.z, gv, dv[6][cc]); dv[7][cc] = fmaf(pb.w, gv, dv[7][cc]);
      }
    }
    __syncwarp();
    // dK += dS^T Q   (Qs holds q * scale; dS carries one factor of scale already, so un-scale Q once)
    *reinterpret_cast<float4*>(Pw + lane * 8) = make_float4(ds[0], ds[1], ds[2], ds[3]);
    *reinterpret_cast<float4*>(Pw + lane * 8 + 4) = make_float4(ds[4], ds[5], ds[6], ds[7]);
    __syncwarp();
#pragma unroll 2
    for (int j = 0; j < 32; ++j) {
      const float4 pa = *reinterpret_cast<const float4*>(Pw + j * 8);
      const float4 pb = *reinterpret_cast<const float4*>(Pw + j * 8 + 4);
#pragma unroll
      for (int cc = 0; cc < NCC; ++cc) {
        const int c = cc * 32 + lane;
        const float qv = Qs[j * KS + (c < D ? c : 0)];
        dk[0][cc] = fmaf(pa.x, qv, dk[0][cc]); dk[1][cc] = fmaf(pa.y, qv, dk[1][cc]);
        dk[2][cc] = fmaf(pa.z, qv, dk[2][cc]); dk[3][cc] = fmaf(pa.w, qv, dk[3][cc]);
        dk[4][cc] = fmaf(pb.x, qv, dk[4][cc]); dk[5][cc] = fmaf(pb.y, qv, dk[5][cc]);
        dk[6][cc] = fmaf(pb.z, qv, dk[6][cc]); dk[7][cc] = fmaf(pb.w, qv, dk[7][cc]);
      }
    }
    __syncwarp();
  }
  const float inv_scale = 1.0f / p.scale;
#pragma unroll
  for (int r = 0; r < 8; ++r) {
    const int t = k0 + r;
    if (t >= p.nk) continue;
    const long long row = base + static_cast<long long>(t) * p.inner;
    __nv_bfloat16* dkp = p.dK + row * p.lddk + p.dk_col0 + head * p.head_stride;
    __nv_bfloat16* dvp = p.dV + row * p.lddv + p.dv_col0 + head * D;
#pragma unroll
    for (int cc = 0; cc < NCC; ++cc) {
      const int c = cc * 32 + lane;
      if (c < D) {
        dkp[c] = __float2bfloat16(dk[r][cc] * inv_scale);
        dvp[c] = __float2bfloat16(dv[r][cc]);
      }
    }
  }
}

template <int D>
static int launch_attention_bwd(const AttnBwdParams& p, bool want_dkv, cudaStream_t stream) {
  const int smem1 = (2 * 32 * (D + 1) + 2 * 4 * D * 8 + 4 * 32 * 8) * 4;
  const int smem2 = (2 * 32 * (D + 1) + 2 * 4 * D * 8 + 4 * 32 * 8 + 64) * 4;
  static unsigned long long attr_devs = 0;
  if (first_use_on_this_device(&attr_devs)) {
    FMC_CUDA_OK(cudaFuncSetAttribute(attention_bwd_dq_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem1));
    FMC_CUDA_OK(cudaFuncSetAttribute(attention_bwd_dkv_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem2));
  }
  FMC_CUDA_OK(launch_k(attention_bwd_dq_kernel<D>, dim3(ceil_div(p.nq, 32), p.heads, p.images), dim3(128), smem1, stream, p));
  int rc = check_launch("attention_bwd_dq_kernel");
  if (rc != FMC_OK || !want_dkv) return rc;
  FMC_CUDA_OK(launch_k(attention_bwd_dkv_kernel<D>, dim3(ceil_div(p.nk, 32), p.heads, p.images), dim3(128), smem2, stream, p));
  return check_launch("attention_bwd_dkv_kernel");
}

// ---------------------------------------------------------------------------------------------------------------
// Temporal self-attention backward for 16-frame sequences (every motion-module attention of the training configs):
// one warp per (sequence, head), everything of the 16 x 16 problem on chip.  The four operand slices (Q, K, V, dO:
// 16 rows x D) are staged in shared memory as bf16 with a 16-byte row pad; phase 1 has lane = (query i, half of the keys)
// and computes S, P, dP, D_i = sum_j P_ij dP_ij (= sum_c dO_ic O_ic, so O is not read) and dS in fp32; phase 2 has
// lane = (row, half of the channels) and forms dQ = dS K, dK = dS^T Q, dV = P^T dO in 20-channel chunks.
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void unpack8f(const uint4& v, float (&f)[8]) {
  f[0] = bf16_lo(v.x); f[1] = bf16_hi(v.x); f[2] = bf16_lo(v.y); f[3] = bf16_hi(v.y);
  f[4] = bf16_lo(v.z); f[5] = bf16_hi(v.z); f[6] = bf16_lo(v.w); f[7] = bf16_hi(v.w);
}

template <int D, int WARPS>
__global__ void __launch_bounds__(WARPS * 32)
temporal16_attn_bwd_kernel(AttnBwdParams p) {
  pdl_launch_dependents();
  pdl_wait();
  constexpr int F = 16, VPR = D / 8, PITCH = D * 2 + 16, TILE = F * PITCH, PS = 17;
  constexpr int WARP_BYTES = 4 * TILE + 2 * F * PS * 4;
  extern __shared__ __align__(16) uint8_t smem_t16[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long w = static_cast<long long>(blockIdx.x) * WARPS + warp;
  if (w >= static_cast<long long>(p.images) * p.heads) return;  // warp-uniform; only __syncwarp below
  const int seq = static_cast<int>(w / p.heads), head = static_cast<int>(w % p.heads);
  const long long row0 = static_cast<long long>(seq / p.inner) * F * p.inner + (seq % p.inner);
  uint8_t* sq = smem_t16 + warp * WARP_BYTES;
  uint8_t* sk = sq + TILE;
  uint8_t* sv = sk + TILE;
  uint8_t* sg = sv + TILE;
  float* sp = reinterpret_cast<float*>(sg + TILE);  // P  [16][17]
  float* sd = sp + F * PS;                          // dS [16][17]
  for (int idx = lane; idx < 4 * F * VPR; idx += 32) {
    const int op = idx / (F * VPR), rem = idx % (F * VPR), t = rem / VPR, v = rem % VPR;
    const long long row = row0 + static_cast<long long>(t) * p.inner;
    const __nv_bfloat16* src = op == 0   ? p.Q + row * p.ldq + p.q_col0 + head * p.head_stride
                               : op == 1 ? p.K + row * p.ldk + p.k_col0 + head * p.head_stride
                               : op == 2 ? p.V + row * p.ldv + p.v_col0 + head * D
                                         : p.dO + row * p.lddo + head * D;
    *reinterpret_cast<uint4*>(sq + op * TILE + t * PITCH + v * 16) = __ldg(reinterpret_cast<const uint4*>(src) + v);
  }
  __syncwarp();
  {
    const int i = lane >> 1, j0 = (lane & 1) * 8;
    float s[8], dp[8];
#pragma unroll
    for (int jj = 0; jj < 8; ++jj) s[jj] = dp[jj] = 0.f;
#pragma unroll 1
    for (int v = 0; v < VPR; ++v) {
      float q8[8], g8[8];
      unpack8f(*reinterpret_cast<const uint4*>(sq + i * PITCH + v * 16), q8);
      unpack8f(*reinterpret_cast<const uint4*>(sg + i * PITCH + v * 16), g8);
#pragma unroll
      for (int jj = 0; jj < 8; ++jj) {
        float k8[8], v8[8];
        unpack8f(*reinterpret_cast<const uint4*>(sk + (j0 + jj) * PITCH + v * 16), k8);
        unpack8f(*reinterpret_cast<const uint4*>(sv + (j0 + jj) * PITCH + v * 16), v8);
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          s[jj] = fmaf(q8[e], k8[e], s[jj]);
          dp[jj] = fmaf(g8[e], v8[e], dp[jj]);
        }
      }
    }
    float mx = -INFINITY;
#pragma unroll
    for (int jj = 0; jj < 8; ++jj) { s[jj] *= p.scale; mx = fmaxf(mx, s[jj]); }
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
    float sum = 0.f;
#pragma unroll
    for (int jj = 0; jj < 8; ++jj) { s[jj] = __expf(s[jj] - mx); sum += s[jj]; }
    sum += __shfl_xor_sync(0xffffffffu, sum, 1);
    const float inv = 1.f / sum;
    float dsum = 0.f;
#pragma unroll
    for (int jj = 0; jj < 8; ++jj) { s[jj] *= inv; dsum = fmaf(s[jj], dp[jj], dsum); }
    dsum += __shfl_xor_sync(0xffffffffu, dsum, 1);
#pragma unroll
    for (int jj = 0; jj < 8; ++jj) {
      sp[i * PS + j0 + jj] = s[jj];
      sd[i * PS + j0 + jj] = s[jj] * (dp[jj] - dsum) * p.scale;
    }
  }
  __syncwarp();
  {
    const int r = lane & 15, ch = lane >> 4;
    const long long row = row0 + static_cast<long long>(r) * p.inner;
    __nv_bfloat16* dq = p.dQ + row * p.lddq + p.dq_col0 + head * p.head_stride;
    __nv_bfloat16* dk = p.dK + row * p.lddk + p.dk_col0 + head * p.head_stride;
    __nv_bfloat16* dv = p.dV + row * p.lddv + p.dv_col0 + head * D;
#pragma unroll 1
    for (int which = 0; which < 3; ++which) {
      const uint8_t* src = which == 0 ? sk : which == 1 ? sq : sg;  // dQ = dS K, dK = dS^T Q, dV = P^T dO
      const float* coef = which == 2 ? sp : sd;
      const int cstride_t = which == 0 ? 1 : PS, cstride_r = which == 0 ? PS : 1;  // coefficient of (own row r, other t)
      __nv_bfloat16* dst = which == 0 ? dq : which == 1 ? dk : dv;
#pragma unroll 1
      for (int cc = 0; cc < D / 40; ++cc) {
        const int c0 = ch * (D / 2) + cc * 20;
        float acc[20];
#pragma unroll
        for (int e = 0; e < 20; ++e) acc[e] = 0.f;
#pragma unroll 4
        for (int t = 0; t < F; ++t) {
          const float a = coef[r * cstride_r + t * cstride_t];
          const uint2* sp2 = reinterpret_cast<const uint2*>(src + t * PITCH + c0 * 2);
#pragma unroll
          for (int v = 0; v < 5; ++v) {
            const uint2 x = sp2[v];
            acc[4 * v] = fmaf(a, bf16_lo(x.x), acc[4 * v]);
            acc[4 * v + 1] = fmaf(a, bf16_hi(x.x), acc[4 * v + 1]);
            acc[4 * v + 2] = fmaf(a, bf16_lo(x.y), acc[4 * v + 2]);
            acc[4 * v + 3] = fmaf(a, bf16_hi(x.y), acc[4 * v + 3]);
          }
        }
#pragma unroll
        for (int v = 0; v < 5; ++v)
          reinterpret_cast<uint2*>(dst + c0)[v] =
              make_uint2(pack_bf16x2(acc[4 * v], acc[4 * v + 1]), pack_bf16x2(acc[4 * v + 2], acc[4 * v + 3]));
      }
    }
  }
}

template <int D, int WARPS>
static int launch_temporal16_attn_bwd(const AttnBwdParams& p, cudaStream_t stream) {
  constexpr int smem = WARPS * (4 * 16 * (D * 2 + 16) + 2 * 16 * 17 * 4);
  static unsigned long long attr_devs = 0;
  if (first_use_on_this_device(&attr_devs))
    FMC_CUDA_OK(cudaFuncSetAttribute(temporal16_attn_bwd_kernel<D, WARPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  const long long warps = static_cast<long long>(p.images) * p.heads;
  FMC_CUDA_OK(launch_k(temporal16_attn_bwd_kernel<D, WARPS>, dim3(static_cast<unsigned>((warps + WARPS - 1) / WARPS)),
                       dim3(WARPS * 32), smem, stream, p));
  return check_launch("temporal16_attn_bwd_kernel");
}

}  // namespace fmc

using namespace fmc;

extern "C" int fmc_transpose_bf16(const void* x, long long ldx, void* out, long long ldo, long long rows, int cols,
                                  void* stream) {
  FMC_REQUIRE(x && out, FMC_ERR_ARG, "fmc_transpose_bf16: null operand");
  if (rows == 0 || cols == 0) return FMC_OK;
  FMC_REQUIRE((rows + 31) / 32 <= 65535, FMC_ERR_SHAPE, "fmc_transpose_bf16: too many rows (%lld)", rows);
  launch_k(transpose_bf16_kernel, dim3(ceil_div(cols, 32), static_cast<unsigned>((rows + 31) / 32)), dim3(256), 0,
           static_cast<cudaStream_t>(stream), static_cast<const __nv_bfloat16*>(x), ldx, static_cast<__nv_bfloat16*>(out), ldo,
           rows, cols);
  return check_launch("transpose_bf16_kernel");
}

extern "C" int fmc_colsum_workspace_floats(long long rows, int cols) {
  return static_cast<int>(((rows + COLSUM_ROWS - 1) / COLSUM_ROWS) * cols);
}

extern "C" int fmc_colsum_f32(const void* x, long long ldx, int x_is_bf16, float* out, float* workspace, long long rows,
                              int cols, int accumulate, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  FMC_REQUIRE(x && out && workspace, FMC_ERR_ARG, "fmc_colsum_f32: null operand");
  if (cols == 0) return FMC_OK;
  const int chunks = static_cast<int>((rows + COLSUM_ROWS - 1) / COLSUM_ROWS);
  FMC_REQUIRE(chunks <= 65535, FMC_ERR_SHAPE, "fmc_colsum_f32: too many rows (%lld)", rows);
  if (chunks > 0) {
    if (x_is_bf16)
      launch_k(colsum_partial_kernel<__nv_bfloat16>, dim3(ceil_div(cols, 256), chunks), dim3(256), 0, stream,
               static_cast<const __nv_bfloat16*>(x), ldx, workspace, rows, cols);
    else
      launch_k(colsum_partial_kernel<float>, dim3(ceil_div(cols, 256), chunks), dim3(256), 0, stream,
               static_cast<const float*>(x), ldx, workspace, rows, cols);
    int rc = check_launch("colsum_partial_kernel");
    if (rc != FMC_OK) return rc;
  }
  launch_k(colsum_final_kernel, dim3(ceil_div(cols, 256)), dim3(256), 0, stream, static_cast<const float*>(workspace), out,
           chunks, cols, accumulate);
  return check_launch("colsum_final_kernel");
}

extern "C" int fmc_layernorm_bwd_blocks(long long rows) {
  return static_cast<int>((rows + LNB_ROWS_PER_BLOCK - 1) / LNB_ROWS_PER_BLOCK);
}

extern "C" int fmc_layernorm_bwd_bf16(const void* x, long long ldx, const void* dy, long long lddy, const float* gamma,
                                      float eps, void* dx, long long lddx, float* param_partials, long long rows, int C,
                                      void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  FMC_REQUIRE(x && dy && gamma && dx, FMC_ERR_ARG, "fmc_layernorm_bwd_bf16: null operand");
  FMC_REQUIRE(C % 8 == 0 && C <= 1280 && ldx % 8 == 0 && lddy % 8 == 0 && lddx % 8 == 0, FMC_ERR_SHAPE,
              "fmc_layernorm_bwd_bf16: C=%d must be a multiple of 8, at most 1280", C);
  if (rows == 0) return FMC_OK;
  // with parameter partials the block count is part of the interface (fmc_layernorm_bwd_blocks); without them the rows
  // per block shrink until the grid covers the SMs several times over
  const bool params = param_partials != nullptr;
  int rpb = LNB_ROWS_PER_BLOCK;
  if (!params)
    while (rpb > 2 * LNB_WARPS && (rows + rpb - 1) / rpb < 8ll * device_sm_count()) rpb >>= 1;
  const int blocks = static_cast<int>((rows + rpb - 1) / rpb);
  const size_t smem = params ? static_cast<size_t>(LNB_WARPS) * 2 * C * sizeof(float) : 0;
  const __nv_bfloat16* xb = static_cast<const __nv_bfloat16*>(x);
  const __nv_bfloat16* db = static_cast<const __nv_bfloat16*>(dy);
  __nv_bfloat16* ob = static_cast<__nv_bfloat16*>(dx);
#define FMC_LNB_LAUNCH(NV, PARAMS, R, SMEM_MAX)                                                                             \
  do {                                                                                                                      \
    static unsigned long long devs = 0;                                                                                     \
    if (PARAMS && first_use_on_this_device(&devs))                                                                          \
      FMC_CUDA_OK(cudaFuncSetAttribute(layernorm_bwd_kernel<NV, PARAMS, R>, cudaFuncAttributeMaxDynamicSharedMemorySize,     \
                                       SMEM_MAX));                                                                          \
    launch_k(layernorm_bwd_kernel<NV, PARAMS, R>, dim3(blocks), dim3(LNB_WARPS * 32), smem, stream, xb, ldx, db, lddy, gamma, \
             eps, ob, lddx, param_partials, rows, C, rpb);                                                                  \
  } while (0)
  if (C <= 512) {
    if (params) FMC_LNB_LAUNCH(2, true, 1, LNB_WARPS * 2 * 512 * 4);
    else FMC_LNB_LAUNCH(2, false, 2, 0);
  } else {
    if (params) FMC_LNB_LAUNCH(5, true, 1, LNB_WARPS * 2 * 1280 * 4);
    else FMC_LNB_LAUNCH(5, false, 1, 0);
  }
#undef FMC_LNB_LAUNCH
  return check_launch("layernorm_bwd_kernel");
}

extern "C" long long fmc_groupnorm_bwd_workspace_floats(int images, int HW, int groups) {
  return 4ll * images * ((HW + GNB_ROWS - 1) / GNB_ROWS) * groups;
}

extern "C" int fmc_groupnorm_bwd_bf16(const void* x, long long ldx, const void* dy, long long lddy, const float* gamma,
                                      const float* beta, float eps, void* dx, long long lddx, float* stats_ws, int images,
                                      int HW, int C, int groups, int silu, const float* rowbias, long long ldrb,
                                      int rowbias_div, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  FMC_REQUIRE(x && dy && gamma && beta && dx && stats_ws, FMC_ERR_ARG, "fmc_groupnorm_bwd_bf16: null operand");
  FMC_REQUIRE(groups > 0 && groups <= 64 && C % groups == 0 && C % 8 == 0 && ldx % 8 == 0 && lddy % 8 == 0 && lddx % 8 == 0,
              FMC_ERR_SHAPE, "fmc_groupnorm_bwd_bf16: C=%d groups=%d", C, groups);
  if (images == 0 || HW == 0) return FMC_OK;
  FMC_REQUIRE(images <= 65535, FMC_ERR_SHAPE, "fmc_groupnorm_bwd_bf16: grid limit (images=%d)", images);
  GnbParams p;
  p.x = static_cast<const __nv_bfloat16*>(x); p.dy = static_cast<const __nv_bfloat16*>(dy);
  p.dx = static_cast<__nv_bfloat16*>(dx);
  p.ldx = ldx; p.lddy = lddy; p.lddx = lddx; p.ldrb = ldrb;
  p.gamma = gamma; p.beta = beta; p.rowbias = rowbias; p.eps = eps;
  p.images = images; p.HW = HW; p.C = C; p.groups = groups; p.silu = silu; p.rb_div = rowbias_div > 0 ? rowbias_div : 1;
  // chunks: enough (image, chunk) blocks for four per SM, but at least GNB_ROWS rows each (the per-thread channel constants
  // and the fold of the partials are paid once per block); the workspace bound ceil(HW / GNB_ROWS) chunks always covers it
  const int max_chunks = (HW + GNB_ROWS - 1) / GNB_ROWS;
  int want = (4 * device_sm_count() + images - 1) / images;
  want = want < 1 ? 1 : (want > max_chunks ? max_chunks : want);
  p.chunk_rows = ((HW + want - 1) / want + 7) / 8 * 8;
  p.chunks = (HW + p.chunk_rows - 1) / p.chunk_rows;
  p.part0 = stats_ws;
  p.part1 = stats_ws + 2ll * images * p.chunks * groups;
  const int nvec = C / 8, VT = nvec < 256 ? nvec : 256, RL = 256 / VT;
  const int smem = RL * C * 2 * static_cast<int>(sizeof(float));
  FMC_REQUIRE(smem <= 96 * 1024, FMC_ERR_SHAPE, "fmc_groupnorm_bwd_bf16: C=%d too wide", C);
  static unsigned long long devs = 0;
  if (first_use_on_this_device(&devs)) {
    FMC_CUDA_OK(cudaFuncSetAttribute(groupnorm_bwd_pass_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
    FMC_CUDA_OK(cudaFuncSetAttribute(groupnorm_bwd_pass_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
  }
  const dim3 grid(p.chunks, images);
  launch_k(groupnorm_bwd_pass_kernel<0>, grid, dim3(256), smem, stream, p);
  int rc = check_launch("groupnorm_bwd_pass_kernel<0>");
  if (rc != FMC_OK) return rc;
  launch_k(groupnorm_bwd_pass_kernel<1>, grid, dim3(256), smem, stream, p);
  rc = check_launch("groupnorm_bwd_pass_kernel<1>");
  if (rc != FMC_OK) return rc;
  launch_k(groupnorm_bwd_pass_kernel<2>, grid, dim3(256), 0, stream, p);
  return check_launch("groupnorm_bwd_pass_kernel<2>");
}

extern "C" int fmc_geglu_fwd_bf16(const void* proj, long long ldp, void* y, long long ldy, long long rows, int H,
                                  void* stream) {
  FMC_REQUIRE(proj && y, FMC_ERR_ARG, "fmc_geglu_fwd_bf16: null operand");
  FMC_REQUIRE(H % 16 == 0 && ldp % 8 == 0 && ldy % 8 == 0, FMC_ERR_SHAPE, "fmc_geglu_fwd_bf16: H=%d must be a multiple of 16", H);
  if (rows == 0) return FMC_OK;
  launch_k(geglu_fwd_kernel, dim3(bblocks(rows * (H / 8))), dim3(256), 0, static_cast<cudaStream_t>(stream),
           static_cast<const __nv_bfloat16*>(proj), ldp, static_cast<__nv_bfloat16*>(y), ldy, rows, H);
  return check_launch("geglu_fwd_kernel");
}

extern "C" int fmc_geglu_bwd_bf16(const void* proj, long long ldp, const void* dy, long long lddy, void* dproj,
                                  long long lddp, long long rows, int H, void* stream) {
  FMC_REQUIRE(proj && dy && dproj, FMC_ERR_ARG, "fmc_geglu_bwd_bf16: null operand");
  FMC_REQUIRE(H % 16 == 0 && ldp % 8 == 0 && lddy % 8 == 0 && lddp % 8 == 0, FMC_ERR_SHAPE,
              "fmc_geglu_bwd_bf16: H=%d must be a multiple of 16", H);
  if (rows == 0) return FMC_OK;
  launch_k(geglu_bwd_kernel, dim3(bblocks(rows * (H / 8))), dim3(256), 0, static_cast<cudaStream_t>(stream),
           static_cast<const __nv_bfloat16*>(proj), ldp, static_cast<const __nv_bfloat16*>(dy), lddy,
           static_cast<__nv_bfloat16*>(dproj), lddp, rows, H);
  return check_launch("geglu_bwd_kernel");
}

extern "C" int fmc_relu_bwd_bf16(const void* y, const void* dy, void* dx, long long n, void* stream) {
  FMC_REQUIRE(y && dy && dx && n % 8 == 0, FMC_ERR_ARG, "fmc_relu_bwd_bf16: bad arguments");
  if (n == 0) return FMC_OK;
  launch_k(relu_bwd_kernel, dim3(bblocks(n / 8)), dim3(256), 0, static_cast<cudaStream_t>(stream),
           static_cast<const __nv_bfloat16*>(y), static_cast<const __nv_bfloat16*>(dy), static_cast<__nv_bfloat16*>(dx), n / 8);
  return check_launch("relu_bwd_kernel");
}

extern "C" int fmc_resize_nearest_bwd_bf16(const void* dy, void* dx, int N, int h, int w, int oh, int ow, int C,
                                           void* stream) {
  FMC_REQUIRE(dy && dx, FMC_ERR_ARG, "fmc_resize_nearest_bwd_bf16: null operand");
  FMC_REQUIRE(C % 8 == 0 && h > 0 && w > 0 && oh % h == 0 && ow % w == 0, FMC_ERR_SHAPE,
              "fmc_resize_nearest_bwd_bf16: %dx%d -> %dx%d is not an integer upsampling", h, w, oh, ow);
  const long long total = static_cast<long long>(N) * h * w * (C / 8);
  if (total == 0) return FMC_OK;
  launch_k(resize_nearest_bwd_kernel, dim3(bblocks(total)), dim3(256), 0, static_cast<cudaStream_t>(stream),
           static_cast<const __nv_bfloat16*>(dy), static_cast<__nv_bfloat16*>(dx), N, h, w, oh / h, ow / w, C);
  return check_launch("resize_nearest_bwd_kernel");
}

extern "C" int fmc_avgpool2_bwd_bf16(const void* dy, void* dx, int N, int h, int w, int C, void* stream) {
  FMC_REQUIRE(dy && dx, FMC_ERR_ARG, "fmc_avgpool2_bwd_bf16: null operand");
  FMC_REQUIRE(C % 8 == 0, FMC_ERR_SHAPE, "fmc_avgpool2_bwd_bf16: C must be a multiple of 8");
  const long long total = static_cast<long long>(N) * h * w * (C / 8);
  if (total == 0) return FMC_OK;
  launch_k(avgpool2_bwd_kernel, dim3(bblocks(total)), dim3(256), 0, static_cast<cudaStream_t>(stream),
           static_cast<const __nv_bfloat16*>(dy), static_cast<__nv_bfloat16*>(dx), N, h, w, C);
  return check_launch("avgpool2_bwd_kernel");
}

namespace fmc {
int attention_bwd_tc(int head_dim, const void* Q, long long ldq, int q_col0, const void* K, long long ldk, int k_col0, const void* V,
                         long long ldv, int v_col0, int head_stride, const void* O, long long ldo, const void* dO,
                         long long lddo, void* dQ, long long lddq, int dq_col0, void* dK, long long lddk, int dk_col0,
                         void* dV, long long lddv, int dv_col0, float* lse, float* dsum, int images, int heads, int n, int nk,
                         int kv_div, int kv_stride, float scale, bool have_lse, cudaStream_t stream);
}
extern "C" int fmc_attention_bwd_bf16(const void* Q, long long ldq, int q_col0, const void* K, long long ldk, int k_col0,
                                      const void* V, long long ldv, int v_col0, int head_stride, const void* O,
                                      long long ldo, const void* dO, long long lddo, void* dQ, long long lddq, int dq_col0,
                                      void* dK, long long lddk, int dk_col0, void* dV, long long lddv, int dv_col0,
                                      float* lse, float* dsum, int lse_given, int images, int heads, int head_dim, int nq,
                                      int nk, int kv_div, int kv_stride, int inner, float scale, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  FMC_REQUIRE(Q && K && V && O && dO && dQ && lse && dsum, FMC_ERR_ARG, "fmc_attention_bwd_bf16: null operand");
  FMC_REQUIRE((dK == nullptr) == (dV == nullptr), FMC_ERR_ARG, "fmc_attention_bwd_bf16: dK and dV go together");
  FMC_REQUIRE(images > 0 && heads > 0 && nq > 0 && nk > 0 && kv_div > 0 && inner > 0, FMC_ERR_ARG,
              "fmc_attention_bwd_bf16: sizes must be positive");
  FMC_REQUIRE(dK == nullptr || (kv_div == 1 && nq == nk && kv_stride == nk), FMC_ERR_ARG,
              "fmc_attention_bwd_bf16: key / value gradients are implemented for self-attention only");
  FMC_REQUIRE(images <= 65535, FMC_ERR_SHAPE, "fmc_attention_bwd_bf16: grid limit (images=%d)", images);
  FMC_REQUIRE(ldk % 8 == 0 && ldv % 8 == 0 && ldq % 8 == 0 && lddo % 8 == 0 && k_col0 % 8 == 0 && v_col0 % 8 == 0 &&
                  q_col0 % 8 == 0 && head_stride % 8 == 0 && head_dim % 8 == 0, FMC_ERR_SHAPE,
              "fmc_attention_bwd_bf16: rows and heads must be 16-byte aligned");
  AttnBwdParams p;
  p.Q = static_cast<const __nv_bfloat16*>(Q); p.K = static_cast<const __nv_bfloat16*>(K);
  p.V = static_cast<const __nv_bfloat16*>(V); p.O = static_cast<const __nv_bfloat16*>(O);
  p.dO = static_cast<const __nv_bfloat16*>(dO);
  p.dQ = static_cast<__nv_bfloat16*>(dQ); p.dK = static_cast<__nv_bfloat16*>(dK); p.dV = static_cast<__nv_bfloat16*>(dV);
  p.ldq = ldq; p.ldk = ldk; p.ldv = ldv; p.ldo = ldo; p.lddo = lddo; p.lddq = lddq; p.lddk = lddk; p.lddv = lddv;
  p.q_col0 = q_col0; p.k_col0 = k_col0; p.v_col0 = v_col0; p.dq_col0 = dq_col0; p.dk_col0 = dk_col0; p.dv_col0 = dv_col0;
  p.head_stride = head_stride; p.lse = lse; p.dsum = dsum;
  p.images = images; p.heads = heads; p.nq = nq; p.nk = nk; p.kv_div = kv_div; p.kv_stride = kv_stride; p.inner = inner;
  p.scale = scale;
  // level-0 spatial self-attention (the bulk of the training step) runs on the tensor cores (attn_bwd_tc.cu);
  // FMC_ATTN_BWD_SIMT=1 keeps the SIMT kernels below for A/B checks
  const char* simt_env = getenv("FMC_ATTN_BWD_SIMT");
  const bool force_simt = simt_env && simt_env[0] == '1';
  const bool aligned8 = ldo % 8 == 0 && lddq % 8 == 0 && lddk % 8 == 0 && lddv % 8 == 0 && dq_col0 % 8 == 0 &&
                        dk_col0 % 8 == 0 && dv_col0 % 8 == 0;
  // 16-frame self-attention (every motion-module attention): one warp per (sequence, head), all on chip
  if (nq == 16 && nk == 16 && dK != nullptr && !force_simt && aligned8) {
    switch (head_dim) {
      case 40: return launch_temporal16_attn_bwd<40, 8>(p, stream);
      case 80: return launch_temporal16_attn_bwd<80, 8>(p, stream);
      case 160: return launch_temporal16_attn_bwd<160, 4>(p, stream);
      default: break;
    }
  }
  // spatial self-attention and (dQ only) text cross-attention
  if (((head_dim == 40 && head_stride == 48) || (head_dim == 80 && head_stride == 80) ||
       (head_dim == 160 && head_stride == 160)) && inner == 1 && !force_simt &&
      ldo % 8 == 0 && lddq % 8 == 0 && dq_col0 % 8 == 0 && (dK == nullptr || aligned8) &&
      static_cast<long long>(images) * nq < 0x7fffffffll)
    return attention_bwd_tc(head_dim, Q, ldq, q_col0, K, ldk, k_col0, V, ldv, v_col0, head_stride, O, ldo, dO, lddo, dQ, lddq,
                            dq_col0, dK, lddk, dk_col0, dV, lddv, dv_col0, lse, dsum, images, heads, nq, nk, kv_div, kv_stride,
                            scale, lse_given != 0 && dK != nullptr, stream);
  switch (head_dim) {
    case 40: return launch_attention_bwd<40>(p, dK != nullptr, stream);
    case 80: return launch_attention_bwd<80>(p, dK != nullptr, stream);
    case 160: return launch_attention_bwd<160>(p, dK != nullptr, stream);
    default: break;
  }
  set_error("fmc_attention_bwd_bf16: head_dim %d not in {40, 80, 160}", head_dim);
  return FMC_ERR_SHAPE;
}
