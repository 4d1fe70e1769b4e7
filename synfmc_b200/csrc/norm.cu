// Normalisation kernels on channels-last bf16 activations (HBM-bound; fp32 statistics).
//   LayerNorm (+ temporal positional encoding + CameraAdapter pose add)   fmc/models/motion_module.py:289,320,355-356
//                                                                          fmc/models/attention_processor.py:257 (x + pose)
//   GroupNorm(32) per frame (+ SiLU)                                        fmc/models/resnet.py:27-37, diffusers ResnetBlock2D,
//                                                                          Transformer2DModel.norm, motion_module.py:217
#include <cuda_bf16.h>

#include "common.cuh"
#include "ptx.cuh"

namespace fmc {

// ------------------------------------------------------------------------------------------------
// LayerNorm: one warp per row, the row lives in registers (C <= 1280 -> at most 5 x 8 values per lane).
// out  = LN(x) * gamma + beta (+ pe[frame])            (bf16)
// out2 = out_fp32 + add                                 (bf16, optional: CameraAdapter input x + pose)
// ------------------------------------------------------------------------------------------------
constexpr int LN_MAX_VEC = 5;  // per lane: 5 vectors of 8 channels -> C <= 1280

__global__ void __launch_bounds__(256)
layernorm_kernel(const __nv_bfloat16* __restrict__ x, long long ldx, const float* __restrict__ gamma,
                 const float* __restrict__ beta, float eps, __nv_bfloat16* __restrict__ out, long long ldo,
                 const float* __restrict__ pe, int F, int HW, const __nv_bfloat16* __restrict__ add, long long ldadd,
                 __nv_bfloat16* __restrict__ out2, long long ldo2, long long rows, int C) {
  const int lane = threadIdx.x & 31;
  const long long row = blockIdx.x * 8ll + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int nvec = C >> 3;
  float v[LN_MAX_VEC][8];
  float sum = 0.f;
  const uint4* xr = reinterpret_cast<const uint4*>(x + row * ldx);
#pragma unroll
  for (int i = 0; i < LN_MAX_VEC; ++i) {
    const int vi = lane + i * 32;
    if (vi < nvec) {
      const uint4 u = __ldg(xr + vi);
      const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        v[i][2 * j] = bf16_lo(w[j]);
        v[i][2 * j + 1] = bf16_hi(w[j]);
        sum += v[i][2 * j] + v[i][2 * j + 1];
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  const float mean = sum / C;
  float sq = 0.f;
#pragma unroll
  for (int i = 0; i < LN_MAX_VEC; ++i) {
    if (lane + i * 32 < nvec) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float d = v[i][j] - mean;
        sq += d * d;
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
  const float rstd = rsqrtf(sq / C + eps);
  const float* per = pe != nullptr ? pe + static_cast<long long>((row / HW) % F) * C : nullptr;
#pragma unroll
  for (int i = 0; i < LN_MAX_VEC; ++i) {
    const int vi = lane + i * 32;
    if (vi < nvec) {
      const int c0 = vi * 8;
      float y[8];
#pragma unroll
      for (int j = 0; j < 8; j += 4) {
        const float4 g = __ldg(reinterpret_cast<const float4*>(gamma + c0 + j));
        const float4 b = __ldg(reinterpret_cast<const float4*>(beta + c0 + j));
        y[j] = (v[i][j] - mean) * rstd * g.x + b.x;
        y[j + 1] = (v[i][j + 1] - mean) * rstd * g.y + b.y;
        y[j + 2] = (v[i][j + 2] - mean) * rstd * g.z + b.z;
        y[j + 3] = (v[i][j + 3] - mean) * rstd * g.w + b.w;
        if (per != nullptr) {
          const float4 e = __ldg(reinterpret_cast<const float4*>(per + c0 + j));
          y[j] += e.x; y[j + 1] += e.y; y[j + 2] += e.z; y[j + 3] += e.w;
        }
      }
      *reinterpret_cast<uint4*>(out + row * ldo + c0) =
          make_uint4(pack_bf16x2(y[0], y[1]), pack_bf16x2(y[2], y[3]), pack_bf16x2(y[4], y[5]), pack_bf16x2(y[6], y[7]));
      if (out2 != nullptr) {
        const uint4 u = __ldg(reinterpret_cast<const uint4*>(add + row * ldadd + c0));
        const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          y[2 * j] += bf16_lo(w[j]);
          y[2 * j + 1] += bf16_hi(w[j]);
        }
        *reinterpret_cast<uint4*>(out2 + row * ldo2 + c0) =
            make_uint4(pack_bf16x2(y[0], y[1]), pack_bf16x2(y[2], y[3]), pack_bf16x2(y[4], y[5]), pack_bf16x2(y[6], y[7]));
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// GroupNorm statistics: x [images, HW, C] -> stats[images, G, 2] = (sum, sum of squares), accumulated with atomics.
// Each thread owns one 8-channel vector position and walks rows; per-channel partials go to shared per-group sums.
// ------------------------------------------------------------------------------------------------
constexpr int GN_ROWS_PER_BLOCK = 64;

__global__ void __launch_bounds__(1024)
groupnorm_stats_kernel(const __nv_bfloat16* __restrict__ x, long long ldx, float* __restrict__ stats, int HW, int C,
                       int G, const float* __restrict__ rowbias, long long ldrb, int rb_div) {
  __shared__ float s_sum[64], s_sq[64];
  const int img = blockIdx.y;
  const int row0 = blockIdx.x * GN_ROWS_PER_BLOCK;
  const int nvec = C >> 3;
  const int cpg = C / G;
  if (threadIdx.x < 64) {
    s_sum[threadIdx.x] = 0.f;
    s_sq[threadIdx.x] = 0.f;
  }
  __syncthreads();
  const int lanes_per_row = nvec;                       // threads cooperating on one row
  const int rows_par = blockDim.x / lanes_per_row;      // rows processed concurrently (>= 1 when C <= 2048)
  if (rows_par > 0) {
    const int vi = threadIdx.x % lanes_per_row;
    const int rsub = threadIdx.x / lanes_per_row;
    if (rsub < rows_par) {
      float a[8], b[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) a[j] = b[j] = 0.f;
      float rb[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) rb[j] = rowbias != nullptr ? __ldg(rowbias + (img / rb_div) * ldrb + vi * 8 + j) : 0.f;
      const int rend = min(row0 + GN_ROWS_PER_BLOCK, HW);
      for (int r = row0 + rsub; r < rend; r += rows_par) {
        const uint4 u = __ldg(reinterpret_cast<const uint4*>(x + (static_cast<long long>(img) * HW + r) * ldx) + vi);
        const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float lo = bf16_lo(w[j]) + rb[2 * j], hi = bf16_hi(w[j]) + rb[2 * j + 1];
          a[2 * j] += lo; b[2 * j] += lo * lo;
          a[2 * j + 1] += hi; b[2 * j + 1] += hi * hi;
        }
      }
      // fold the 8 channels into (at most two) groups
      int g_prev = (vi * 8) / cpg;
      float sa = 0.f, sb = 0.f;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int g = (vi * 8 + j) / cpg;
        if (g != g_prev) {
          atomicAdd(&s_sum[g_prev], sa);
          atomicAdd(&s_sq[g_prev], sb);
          sa = sb = 0.f;
          g_prev = g;
        }
        sa += a[j];
        sb += b[j];
      }
      atomicAdd(&s_sum[g_prev], sa);
      atomicAdd(&s_sq[g_prev], sb);
    }
  }
  __syncthreads();
  if (threadIdx.x < G) {
    atomicAdd(&stats[(static_cast<long long>(img) * G + threadIdx.x) * 2], s_sum[threadIdx.x]);
    atomicAdd(&stats[(static_cast<long long>(img) * G + threadIdx.x) * 2 + 1], s_sq[threadIdx.x]);
  }
}

// y = (x (+ rowbias) - mean) * rstd * gamma + beta, optional SiLU.  One thread per 8-channel vector.
__global__ void __launch_bounds__(256)
groupnorm_apply_kernel(const __nv_bfloat16* __restrict__ x, long long ldx, const float* __restrict__ stats,
                       const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
                       __nv_bfloat16* __restrict__ out, long long ldo, long long rows, int HW, int C, int G, int silu,
                       const float* __restrict__ rowbias, long long ldrb, int rb_div) {
  const int nvec = C >> 3;
  const long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (idx >= rows * nvec) return;
  const long long row = idx / nvec;
  const int vi = static_cast<int>(idx % nvec);
  const int img = static_cast<int>(row / HW);
  const int cpg = C / G;
  const float inv_n = 1.0f / (static_cast<float>(HW) * cpg);
  const uint4 u = __ldg(reinterpret_cast<const uint4*>(x + row * ldx) + vi);
  const uint32_t w[4] = {u.x, u.y, u.z, u.w};
  float y[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int c = vi * 8 + j;
    const int g = c / cpg;
    const float s = __ldg(stats + (static_cast<long long>(img) * G + g) * 2);
    const float ss = __ldg(stats + (static_cast<long long>(img) * G + g) * 2 + 1);
    const float mean = s * inv_n;
    const float var = fmaxf(ss * inv_n - mean * mean, 0.f);
    const float rstd = rsqrtf(var + eps);
    float xv = (j & 1) ? bf16_hi(w[j >> 1]) : bf16_lo(w[j >> 1]);
    if (rowbias != nullptr) xv += __ldg(rowbias + (img / rb_div) * ldrb + c);
    float t = (xv - mean) * rstd * __ldg(gamma + c) + __ldg(beta + c);
    if (silu) t = t / (1.0f + __expf(-t));
    y[j] = t;
  }
  *reinterpret_cast<uint4*>(out + row * ldo + vi * 8) =
      make_uint4(pack_bf16x2(y[0], y[1]), pack_bf16x2(y[2], y[3]), pack_bf16x2(y[4], y[5]), pack_bf16x2(y[6], y[7]));
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
  const unsigned grid = static_cast<unsigned>((rows + 7) / 8);
  layernorm_kernel<<<grid, 256, 0, stream>>>(static_cast<const __nv_bfloat16*>(x), ldx, gamma, beta, eps,
                                             static_cast<__nv_bfloat16*>(out), ldo, pe, F > 0 ? F : 1, HW > 0 ? HW : 1,
                                             static_cast<const __nv_bfloat16*>(add), ldadd,
                                             static_cast<__nv_bfloat16*>(out2), ldo2, rows, C);
  return check_launch("layernorm_kernel");
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
  FMC_CUDA_OK(cudaMemsetAsync(stats_ws, 0, sizeof(float) * 2 * groups * images, stream));
  dim3 grid(ceil_div(HW, GN_ROWS_PER_BLOCK), images);
  const int nvec = C / 8;
  const int stat_threads = nvec * (nvec >= 512 ? 1 : 512 / nvec);  // whole rows per pass, <= 1024 threads
  groupnorm_stats_kernel<<<grid, stat_threads, 0, stream>>>(static_cast<const __nv_bfloat16*>(x), ldx, stats_ws, HW, C, groups,
                                                   rowbias, ldrb, rowbias_div > 0 ? rowbias_div : 1);
  int rc = check_launch("groupnorm_stats_kernel");
  if (rc != FMC_OK) return rc;
  const long long rows = static_cast<long long>(images) * HW;
  const long long n = rows * (C / 8);
  groupnorm_apply_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, stream>>>(
      static_cast<const __nv_bfloat16*>(x), ldx, stats_ws, gamma, beta, eps, static_cast<__nv_bfloat16*>(out), ldo, rows,
      HW, C, groups, silu, rowbias, ldrb, rowbias_div > 0 ? rowbias_div : 1);
  return check_launch("groupnorm_apply_kernel");
}
