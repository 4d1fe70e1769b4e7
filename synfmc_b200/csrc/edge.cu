// Kernels of the pipeline edges: the VAE of `decode_latents` / `vae.encode` and the CLIP text encoder of
// `_encode_prompt` (fmc/pipelines/pipeline_animation.py:465-567, train_cam_ctrl.py:544-561).  The contractions of both
// models run on the GEMM / convolution / GroupNorm / LayerNorm kernels of the denoising path; this file holds what they
// need beyond those: the row softmax of the VAE's single-head 512-wide attention (scores come from the GEMM in fp32), the
// small causal multi-head attention of the text encoder (77 tokens, 12 heads of 64), quick-GELU, the token + position
// embedding gather, the posterior sample of the encoder and the decoder's output conversion.  All memory-bound or tiny.
#include <cuda_bf16.h>

#include "common.cuh"
#include "ptx.cuh"

namespace fmc {

template <typename T> __device__ __forceinline__ float ld_act(const T* p);
template <> __device__ __forceinline__ float ld_act<float>(const float* p) { return *p; }
template <> __device__ __forceinline__ float ld_act<__nv_bfloat16>(const __nv_bfloat16* p) { return __bfloat162float(*p); }
template <typename T> __device__ __forceinline__ void st_act(T* p, float v);
template <> __device__ __forceinline__ void st_act<float>(float* p, float v) { *p = v; }
template <> __device__ __forceinline__ void st_act<__nv_bfloat16>(__nv_bfloat16* p, float v) { *p = __float2bfloat16(v); }
__device__ __forceinline__ void st_act4(float* p, float a, float b, float c, float d) {
  *reinterpret_cast<float4*>(p) = make_float4(a, b, c, d);
}
__device__ __forceinline__ void st_act4(__nv_bfloat16* p, float a, float b, float c, float d) {
  *reinterpret_cast<uint2*>(p) = make_uint2(pack_bf16x2(a, b), pack_bf16x2(c, d));
}

// ---------------------------------------------------------------------------------------------------------------
// out[r, :] = softmax(scale * s[r, :]) over n columns (fp32 scores in, bf16 or fp32 probabilities out); one block per
// row, the row held in registers (n <= 4096) so the scores are read once.
// ---------------------------------------------------------------------------------------------------------------
constexpr int SMX_THREADS = 256;
constexpr int SMX_MAX_N = 4096;
template <typename T>
__global__ void __launch_bounds__(SMX_THREADS)
softmax_rows_kernel(const float* __restrict__ s, long long lds, T* __restrict__ out, long long ldo, int n, float scale) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ float red[SMX_THREADS / 32];
  __shared__ float bc;
  const float* row = s + static_cast<long long>(blockIdx.x) * lds;
  T* orow = out + static_cast<long long>(blockIdx.x) * ldo;
  const int nvec = n >> 2;
  const float c = scale * 1.4426950408889634f;
  float4 v[SMX_MAX_N / 4 / SMX_THREADS];
  float mx = -INFINITY;
#pragma unroll
  for (int i = 0; i < SMX_MAX_N / 4 / SMX_THREADS; ++i) {
    const int vi = i * SMX_THREADS + threadIdx.x;
    if (vi < nvec) {
      v[i] = __ldg(reinterpret_cast<const float4*>(row) + vi);
      mx = fmaxf(mx, fmaxf(fmaxf(v[i].x, v[i].y), fmaxf(v[i].z, v[i].w)));
    }
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  auto block_reduce = [&](float x, bool is_max) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float y = __shfl_xor_sync(0xffffffffu, x, o);
      x = is_max ? fmaxf(x, y) : x + y;
    }
    if (lane == 0) red[warp] = x;
    __syncthreads();
    if (threadIdx.x == 0) {
      float t = red[0];
      for (int w = 1; w < SMX_THREADS / 32; ++w) t = is_max ? fmaxf(t, red[w]) : t + red[w];
      bc = t;
    }
    __syncthreads();
    const float r = bc;
    __syncthreads();
    return r;
  };
  // the maximum is taken on the raw scores; a negative scale would need the minimum -- scales are positive (d^-1/2)
  mx = block_reduce(mx, true);
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < SMX_MAX_N / 4 / SMX_THREADS; ++i) {
    const int vi = i * SMX_THREADS + threadIdx.x;
    if (vi < nvec) {
      v[i].x = exp2f((v[i].x - mx) * c); v[i].y = exp2f((v[i].y - mx) * c);
      v[i].z = exp2f((v[i].z - mx) * c); v[i].w = exp2f((v[i].w - mx) * c);
      sum += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    }
  }
  sum = block_reduce(sum, false);
  const float inv = 1.0f / sum;
#pragma unroll
  for (int i = 0; i < SMX_MAX_N / 4 / SMX_THREADS; ++i) {
    const int vi = i * SMX_THREADS + threadIdx.x;
    if (vi < nvec) {
      st_act4(orow + vi * 4, v[i].x * inv, v[i].y * inv, v[i].z * inv, v[i].w * inv);
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Small multi-head self-attention (sequence length <= 128, head_dim D in {32, 64, 96, 128}), optionally causal: one block
// per (sequence, head); K and V of the head in shared memory as fp32; a warp per query: lane = key for the scores, lane =
// channel for P V.  q / k / v are column ranges of one [tokens, ld] buffer (the fused projection), heads packed D apart.
// ---------------------------------------------------------------------------------------------------------------
template <typename T, int D>
__global__ void __launch_bounds__(128)
small_mha_kernel(const T* __restrict__ qkv, long long ld, int q_col0, int k_col0, int v_col0, T* __restrict__ out, long long ldo,
                 int Tn, int heads, float scale, int causal) {
  pdl_launch_dependents();
  pdl_wait();
  extern __shared__ float mha_smem[];
  float* Ks = mha_smem;                 // [Tn][D + 1]
  float* Vs = Ks + Tn * (D + 1);        // [Tn][D]
  float* Qs = Vs + Tn * D;              // [4][D]
  float* Ps = Qs + 4 * D;               // [4][128]
  const int seq = blockIdx.x / heads, head = blockIdx.x % heads;
  const long long row0 = static_cast<long long>(seq) * Tn;
  for (int i = threadIdx.x; i < Tn * D; i += 128) {
    const int t = i / D, c = i % D;
    Ks[t * (D + 1) + c] = ld_act(qkv + (row0 + t) * ld + k_col0 + head * D + c);
    Vs[t * D + c] = ld_act(qkv + (row0 + t) * ld + v_col0 + head * D + c);
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* q = Qs + warp * D;
  float* p = Ps + warp * 128;
  for (int i = warp; i < Tn; i += 4) {
    for (int c = lane; c < D; c += 32) q[c] = ld_act(qkv + (row0 + i) * ld + q_col0 + head * D + c);
    __syncwarp();
    const int limit = causal ? i + 1 : Tn;  // keys 0 .. limit-1
    float s[4];
    float mx = -INFINITY;
#pragma unroll
    for (int jj = 0; jj < 4; ++jj) {
      const int j = jj * 32 + lane;
      float acc = -INFINITY;
      if (j < limit) {
        acc = 0.f;
#pragma unroll 8
        for (int c = 0; c < D; ++c) acc = fmaf(q[c], Ks[j * (D + 1) + c], acc);
        acc *= scale;
      }
      s[jj] = acc;
      mx = fmaxf(mx, acc);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    float sum = 0.f;
#pragma unroll
    for (int jj = 0; jj < 4; ++jj) {
      s[jj] = s[jj] == -INFINITY ? 0.f : __expf(s[jj] - mx);
      sum += s[jj];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    const float inv = 1.0f / sum;
#pragma unroll
    for (int jj = 0; jj < 4; ++jj) p[jj * 32 + lane] = s[jj] * inv;
    __syncwarp();
    for (int c = lane; c < D; c += 32) {
      float acc = 0.f;
      for (int j = 0; j < limit; ++j) acc = fmaf(p[j], Vs[j * D + c], acc);
      st_act(out + (row0 + i) * ldo + head * D + c, acc);
    }
    __syncwarp();
  }
}

// y = x * sigmoid(1.702 x)  (transformers `quick_gelu`, CLIP's MLP activation)
template <typename T>
__global__ void __launch_bounds__(256)
quick_gelu_kernel(const T* __restrict__ x, long long ldx, T* __restrict__ out, long long ldo, long long rows, int cols) {
  pdl_launch_dependents();
  pdl_wait();
  const long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (idx >= rows * cols) return;
  const long long r = idx / cols;
  const int c = static_cast<int>(idx % cols);
  const float v = ld_act(x + r * ldx + c);
  st_act(out + r * ldo + c, v / (1.0f + __expf(-1.702f * v)));
}

// out[b t, :] = token_table[ids[b, t], :] + position_table[t, :]   (CLIPTextEmbeddings.forward)
template <typename T>
__global__ void __launch_bounds__(256)
embed_tokens_kernel(const long long* __restrict__ ids, const float* __restrict__ tok, const float* __restrict__ pos,
                    T* __restrict__ out, long long ldo, long long tokens, int Tn, int C, int vocab) {
  pdl_launch_dependents();
  pdl_wait();
  const long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (idx >= tokens * C) return;
  const long long r = idx / C;
  const int c = static_cast<int>(idx % C);
  long long id = ids[r];
  id = id < 0 ? 0 : (id >= vocab ? vocab - 1 : id);
  st_act(out + r * ldo + c, __ldg(tok + id * C + c) + __ldg(pos + (r % Tn) * C + c));
}

// DiagonalGaussianDistribution.sample (diffusers): moments rows [n, 2 z] channels-last (mean | logvar) -> latents
// [N, z, h, w] fp32 = (mean + exp(0.5 clamp(logvar, -30, 20)) * noise) * out_scale; noise in the output layout (or NULL: mode)
template <typename T>
__global__ void __launch_bounds__(256)
vae_sample_kernel(const T* __restrict__ moments, long long ldm, const float* __restrict__ noise, float* __restrict__ out, int N,
                  int z, long long HW, float out_scale) {
  pdl_launch_dependents();
  pdl_wait();
  const long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  const long long total = static_cast<long long>(N) * z * HW;
  if (idx >= total) return;
  const long long hw = idx % HW;
  const int c = static_cast<int>((idx / HW) % z);
  const long long n = idx / (HW * z);
  const T* m = moments + (n * HW + hw) * ldm;
  const float mean = ld_act(m + c);
  float v = mean;
  if (noise != nullptr) {
    const float logvar = fminf(fmaxf(ld_act(m + z + c), -30.0f), 20.0f);
    v = fmaf(expf(0.5f * logvar), noise[idx], mean);
  }
  out[idx] = v * out_scale;
}

// decode_latents tail: channels-last image rows [(b f) HW, ldc] -> video [b, C, f, H W] fp32 = clamp(x * mul + add, lo, hi)
template <typename T>
__global__ void __launch_bounds__(256)
cl_to_video_kernel(const T* __restrict__ x, long long ldc, float* __restrict__ out, int B, int C, int F, long long HW, float mul,
                   float add, float lo, float hi) {
  pdl_launch_dependents();
  pdl_wait();
  const long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  const long long total = static_cast<long long>(B) * C * F * HW;
  if (idx >= total) return;
  const long long hw = idx % HW;
  long long t = idx / HW;
  const int f = static_cast<int>(t % F);
  t /= F;
  const int c = static_cast<int>(t % C);
  const int b = static_cast<int>(t / C);
  const float v = ld_act(x + ((static_cast<long long>(b) * F + f) * HW + hw) * ldc + c);
  out[idx] = fminf(fmaxf(fmaf(v, mul, add), lo), hi);
}

static unsigned blocks_of(long long n) { return static_cast<unsigned>((n + 255) / 256); }

template <typename T, int D>
static int launch_small_mha(const void* qkv, long long ld, int q_col0, int k_col0, int v_col0, void* out, long long ldo, int seqs,
                            int Tn, int heads, float scale, int causal, cudaStream_t stream) {
  const int smem = (Tn * (D + 1) + Tn * D + 4 * D + 4 * 128) * static_cast<int>(sizeof(float));
  static unsigned long long devs = 0;
  if (first_use_on_this_device(&devs))
    FMC_CUDA_OK(cudaFuncSetAttribute(small_mha_kernel<T, D>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
  FMC_CUDA_OK(launch_k(small_mha_kernel<T, D>, dim3(seqs * heads), dim3(128), smem, stream, static_cast<const T*>(qkv), ld,
                       q_col0, k_col0, v_col0, static_cast<T*>(out), ldo, Tn, heads, scale, causal));
  return check_launch("small_mha_kernel");
}

}  // namespace fmc

using namespace fmc;

extern "C" int fmc_softmax_rows(const float* scores, long long lds, void* out, long long ldo, int out_is_f32, long long rows,
                                int n, float scale, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  FMC_REQUIRE(scores && out, FMC_ERR_ARG, "fmc_softmax_rows: null operand");
  FMC_REQUIRE(n > 0 && n % 4 == 0 && n <= SMX_MAX_N && lds % 4 == 0 && ldo % 4 == 0 && scale > 0.f, FMC_ERR_SHAPE,
              "fmc_softmax_rows: n=%d must be a multiple of 4, at most %d (scale > 0)", n, SMX_MAX_N);
  if (rows == 0) return FMC_OK;
  FMC_REQUIRE(rows <= 0x7fffffffll, FMC_ERR_SHAPE, "fmc_softmax_rows: too many rows (%lld)", rows);
  if (out_is_f32)
    launch_k(softmax_rows_kernel<float>, dim3(static_cast<unsigned>(rows)), dim3(SMX_THREADS), 0, stream, scores, lds,
             static_cast<float*>(out), ldo, n, scale);
  else
    launch_k(softmax_rows_kernel<__nv_bfloat16>, dim3(static_cast<unsigned>(rows)), dim3(SMX_THREADS), 0, stream, scores, lds,
             static_cast<__nv_bfloat16*>(out), ldo, n, scale);
  return check_launch("softmax_rows_kernel");
}

extern "C" int fmc_small_mha(const void* qkv, long long ld, int q_col0, int k_col0, int v_col0, void* out, long long ldo,
                             int is_f32, int seqs, int tokens_per_seq, int heads, int head_dim, float scale, int causal,
                             void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  FMC_REQUIRE(qkv && out, FMC_ERR_ARG, "fmc_small_mha: null operand");
  FMC_REQUIRE(tokens_per_seq > 0 && tokens_per_seq <= 128 && heads > 0, FMC_ERR_SHAPE,
              "fmc_small_mha: %d tokens per sequence (at most 128)", tokens_per_seq);
  if (seqs == 0) return FMC_OK;
#define FMC_MHA_CASE(DD)                                                                                                     \
  case DD:                                                                                                                   \
    return is_f32 ? launch_small_mha<float, DD>(qkv, ld, q_col0, k_col0, v_col0, out, ldo, seqs, tokens_per_seq, heads, scale, \
                                                causal, stream)                                                              \
                  : launch_small_mha<__nv_bfloat16, DD>(qkv, ld, q_col0, k_col0, v_col0, out, ldo, seqs, tokens_per_seq, heads, \
                                                        scale, causal, stream);
  switch (head_dim) {
    FMC_MHA_CASE(32)
    FMC_MHA_CASE(64)
    FMC_MHA_CASE(96)
    FMC_MHA_CASE(128)
    default: break;
  }
#undef FMC_MHA_CASE
  set_error("fmc_small_mha: head_dim %d not in {32, 64, 96, 128}", head_dim);
  return FMC_ERR_SHAPE;
}

extern "C" int fmc_quick_gelu(const void* x, long long ldx, void* out, long long ldo, int is_f32, long long rows, int cols,
                              void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  FMC_REQUIRE(x && out, FMC_ERR_ARG, "fmc_quick_gelu: null operand");
  const long long total = rows * cols;
  if (total == 0) return FMC_OK;
  if (is_f32)
    launch_k(quick_gelu_kernel<float>, dim3(blocks_of(total)), dim3(256), 0, stream, static_cast<const float*>(x), ldx,
             static_cast<float*>(out), ldo, rows, cols);
  else
    launch_k(quick_gelu_kernel<__nv_bfloat16>, dim3(blocks_of(total)), dim3(256), 0, stream,
             static_cast<const __nv_bfloat16*>(x), ldx, static_cast<__nv_bfloat16*>(out), ldo, rows, cols);
  return check_launch("quick_gelu_kernel");
}

extern "C" int fmc_embed_tokens(const long long* ids, const float* token_table, const float* position_table, void* out,
                                long long ldo, int out_is_f32, long long tokens, int tokens_per_seq, int C, int vocab,
                                void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  FMC_REQUIRE(ids && token_table && position_table && out, FMC_ERR_ARG, "fmc_embed_tokens: null operand");
  FMC_REQUIRE(tokens_per_seq > 0 && C > 0 && vocab > 0, FMC_ERR_SHAPE, "fmc_embed_tokens: sizes must be positive");
  const long long total = tokens * C;
  if (total == 0) return FMC_OK;
  if (out_is_f32)
    launch_k(embed_tokens_kernel<float>, dim3(blocks_of(total)), dim3(256), 0, stream, ids, token_table, position_table,
             static_cast<float*>(out), ldo, tokens, tokens_per_seq, C, vocab);
  else
    launch_k(embed_tokens_kernel<__nv_bfloat16>, dim3(blocks_of(total)), dim3(256), 0, stream, ids, token_table, position_table,
             static_cast<__nv_bfloat16*>(out), ldo, tokens, tokens_per_seq, C, vocab);
  return check_launch("embed_tokens_kernel");
}

extern "C" int fmc_vae_sample_f32(const void* moments, long long ldm, int is_f32, const float* noise, float* out, int N, int z,
                                  long long HW, float out_scale, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  FMC_REQUIRE(moments && out && ldm >= 2 * z, FMC_ERR_ARG, "fmc_vae_sample_f32: bad arguments");
  const long long total = static_cast<long long>(N) * z * HW;
  if (total == 0) return FMC_OK;
  if (is_f32)
    launch_k(vae_sample_kernel<float>, dim3(blocks_of(total)), dim3(256), 0, stream, static_cast<const float*>(moments), ldm,
             noise, out, N, z, HW, out_scale);
  else
    launch_k(vae_sample_kernel<__nv_bfloat16>, dim3(blocks_of(total)), dim3(256), 0, stream,
             static_cast<const __nv_bfloat16*>(moments), ldm, noise, out, N, z, HW, out_scale);
  return check_launch("vae_sample_kernel");
}

extern "C" int fmc_cl_to_video_f32(const void* x, long long ldc, int is_f32, float* out, int B, int C, int F, long long HW,
                                   float mul, float add, float lo, float hi, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  FMC_REQUIRE(x && out && ldc >= C, FMC_ERR_ARG, "fmc_cl_to_video_f32: bad arguments");
  const long long total = static_cast<long long>(B) * C * F * HW;
  if (total == 0) return FMC_OK;
  if (is_f32)
    launch_k(cl_to_video_kernel<float>, dim3(blocks_of(total)), dim3(256), 0, stream, static_cast<const float*>(x), ldc, out, B,
             C, F, HW, mul, add, lo, hi);
  else
    launch_k(cl_to_video_kernel<__nv_bfloat16>, dim3(blocks_of(total)), dim3(256), 0, stream,
             static_cast<const __nv_bfloat16*>(x), ldc, out, B, C, F, HW, mul, add, lo, hi);
  return check_launch("cl_to_video_kernel");
}
