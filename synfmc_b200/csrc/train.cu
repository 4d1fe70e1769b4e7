// Training-step tail of the FMC trainers on flat fp32 buffers (train_cam_ctrl.py:647-665,
// train_cam_obj_ctrl.py:843-862): after the gradient all-reduce the reference runs THREE passes over the trainable set
// (218 M parameters for CMC: scaler.unscale_, clip_grad_norm_, AdamW.step).  Here: one deterministic sum-of-squares pass
// (two stages, no atomics), a one-block finalize that leaves the clip coefficient and the non-finite flag ON THE DEVICE
// (no host round trip, so the step can sit in a CUDA graph), and one fused unscale * clip * AdamW pass.  HBM bound:
// 4 B (g) for the norm pass, 16 B read + 12 B written per parameter for the update.
#include "common.cuh"
#include "ptx.cuh"

namespace fmc {

constexpr int SSQ_THREADS = 256;
constexpr int SSQ_MAX_BLOCKS = 1024;

// partial[b] = sum of g^2 over block b's grid-stride slice, bad[b] = 1 if any element is inf / nan
__global__ void __launch_bounds__(SSQ_THREADS)
grad_sumsq_kernel(const float* __restrict__ g, long long n, float* __restrict__ partial, float* __restrict__ bad) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ float red[SSQ_THREADS / 32];
  __shared__ int red_bad[SSQ_THREADS / 32];
  const long long n4 = n >> 2;
  float acc = 0.f;
  int nonfinite = 0;
  for (long long i = blockIdx.x * static_cast<long long>(SSQ_THREADS) + threadIdx.x; i < n4;
       i += static_cast<long long>(gridDim.x) * SSQ_THREADS) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(g) + i);
    acc += (v.x * v.x + v.y * v.y) + (v.z * v.z + v.w * v.w);
    nonfinite |= !(isfinite(v.x) && isfinite(v.y) && isfinite(v.z) && isfinite(v.w));
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {  // tail of a length that is not a multiple of 4
    const float v = g[(n4 << 2) + threadIdx.x];
    acc += v * v;
    nonfinite |= !isfinite(v);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    acc += __shfl_xor_sync(0xffffffffu, acc, o);
    nonfinite |= __shfl_xor_sync(0xffffffffu, nonfinite, o);
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) {
    red[warp] = acc;
    red_bad[warp] = nonfinite;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    int b = 0;
    for (int w = 0; w < SSQ_THREADS / 32; ++w) {
      t += red[w];
      b |= red_bad[w];
    }
    partial[blockIdx.x] = t;
    bad[blockIdx.x] = b ? 1.f : 0.f;
  }
}

// state[0] = total norm of the UN-scaled gradients, state[1] = inv_scale * min(1, max_norm / (norm + 1e-6)) (the factor
// the update multiplies every gradient with: GradScaler.unscale_ followed by clip_grad_norm_), state[2] = found_inf
__global__ void __launch_bounds__(256)
grad_norm_finalize_kernel(const float* __restrict__ partial, const float* __restrict__ bad, int nblocks, float inv_scale,
                          float max_norm, float* __restrict__ state) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ double red[256];
  __shared__ int red_bad[256];
  double acc = 0.0;
  int b = 0;
  for (int i = threadIdx.x; i < nblocks; i += 256) {  // fixed assignment, fixed tree below: deterministic
    acc += static_cast<double>(partial[i]);
    b |= bad[i] != 0.f;
  }
  red[threadIdx.x] = acc;
  red_bad[threadIdx.x] = b;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (threadIdx.x < s) {
      red[threadIdx.x] += red[threadIdx.x + s];
      red_bad[threadIdx.x] |= red_bad[threadIdx.x + s];
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const float norm = inv_scale * static_cast<float>(sqrt(red[0]));
    const bool found_inf = red_bad[0] != 0 || !isfinite(norm);
    float clip = max_norm > 0.f ? max_norm / (norm + 1e-6f) : 1.0f;
    clip = fminf(clip, 1.0f);
    state[0] = norm;
    state[1] = inv_scale * clip;
    state[2] = found_inf ? 1.f : 0.f;
    if (!found_inf) state[3] += 1.f;  // optimizer steps taken so far (fmc_adamw_step_f32 with step = 0 reads it)
  }
}

// torch.optim.AdamW (no amsgrad) on g * state[1]; the whole step is skipped when state[2] != 0 (GradScaler.step)
__global__ void __launch_bounds__(256)
adamw_step_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                  long long n, float lr, float beta1, float beta2, float eps, float weight_decay, float bc1, float rsqrt_bc2,
                  const float* __restrict__ state, int device_step) {
  pdl_launch_dependents();
  pdl_wait();
  const float coef = state != nullptr ? __ldg(state + 1) : 1.0f;
  if (state != nullptr && __ldg(state + 2) != 0.f) return;
  if (device_step) {  // step count (and, for lr < 0, the learning rate) live on the device: the call can sit in a CUDA graph
    const float t = __ldg(state + 3);
    bc1 = 1.0f - exp2f(t * log2f(beta1));
    rsqrt_bc2 = rsqrtf(1.0f - exp2f(t * log2f(beta2)));
    if (lr < 0.f) lr = __ldg(state + 4);
  }
  const long long i4 = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  const long long base = i4 << 2;
  if (base >= n) return;
  const float decay = 1.0f - lr * weight_decay;
  const float step = lr / bc1;
  if (base + 4 <= n) {
    float4 pv = *(reinterpret_cast<float4*>(p) + i4);
    const float4 gv = __ldg(reinterpret_cast<const float4*>(g) + i4);
    float4 mv = *(reinterpret_cast<float4*>(m) + i4);
    float4 vv = *(reinterpret_cast<float4*>(v) + i4);
    float pp[4] = {pv.x, pv.y, pv.z, pv.w}, gg[4] = {gv.x, gv.y, gv.z, gv.w};
    float mm[4] = {mv.x, mv.y, mv.z, mv.w}, vq[4] = {vv.x, vv.y, vv.z, vv.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float gr = gg[j] * coef;
      pp[j] *= decay;
      mm[j] = beta1 * mm[j] + (1.0f - beta1) * gr;      // exp_avg.lerp_(grad, 1 - beta1)
      vq[j] = beta2 * vq[j] + (1.0f - beta2) * gr * gr;  // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, 1 - beta2)
      const float denom = sqrtf(vq[j]) * rsqrt_bc2 + eps;
      pp[j] -= step * (mm[j] / denom);
    }
    *(reinterpret_cast<float4*>(p) + i4) = make_float4(pp[0], pp[1], pp[2], pp[3]);
    *(reinterpret_cast<float4*>(m) + i4) = make_float4(mm[0], mm[1], mm[2], mm[3]);
    *(reinterpret_cast<float4*>(v) + i4) = make_float4(vq[0], vq[1], vq[2], vq[3]);
  } else {
    for (long long i = base; i < n; ++i) {
      const float gr = g[i] * coef;
      float pp = p[i] * decay;
      const float mm = beta1 * m[i] + (1.0f - beta1) * gr;
      const float vq = beta2 * v[i] + (1.0f - beta2) * gr * gr;
      pp -= step * (mm / (sqrtf(vq) * rsqrt_bc2 + eps));
      p[i] = pp; m[i] = mm; v[i] = vq;
    }
  }
}

}  // namespace fmc

using namespace fmc;

extern "C" int fmc_grad_norm_workspace_floats(void) { return 2 * SSQ_MAX_BLOCKS; }

extern "C" int fmc_grad_norm_f32(const float* grad, long long n, float inv_scale, float max_norm, float* workspace,
                                 float* state, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  FMC_REQUIRE(grad && workspace && state, FMC_ERR_ARG, "fmc_grad_norm_f32: null operand");
  FMC_REQUIRE((reinterpret_cast<uintptr_t>(grad) & 15) == 0, FMC_ERR_SHAPE, "fmc_grad_norm_f32: grad not 16-byte aligned");
  FMC_REQUIRE(n >= 0 && inv_scale > 0.f, FMC_ERR_ARG, "fmc_grad_norm_f32: bad arguments");
  long long want = (n / 4 + SSQ_THREADS * 8 - 1) / (SSQ_THREADS * 8);
  const int blocks = static_cast<int>(want < 1 ? 1 : (want > SSQ_MAX_BLOCKS ? SSQ_MAX_BLOCKS : want));
  launch_k(grad_sumsq_kernel, dim3(blocks), dim3(SSQ_THREADS), 0, stream, grad, n, workspace, workspace + SSQ_MAX_BLOCKS);
  int rc = check_launch("grad_sumsq_kernel");
  if (rc != FMC_OK) return rc;
  launch_k(grad_norm_finalize_kernel, dim3(1), dim3(256), 0, stream, static_cast<const float*>(workspace),
           static_cast<const float*>(workspace + SSQ_MAX_BLOCKS), blocks, inv_scale, max_norm, state);
  return check_launch("grad_norm_finalize_kernel");
}

extern "C" int fmc_adamw_step_f32(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, long long n, float lr,
                                  float beta1, float beta2, float eps, float weight_decay, int step, const float* state,
                                  void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  FMC_REQUIRE(param && grad && exp_avg && exp_avg_sq, FMC_ERR_ARG, "fmc_adamw_step_f32: null operand");
  FMC_REQUIRE(((reinterpret_cast<uintptr_t>(param) | reinterpret_cast<uintptr_t>(grad) | reinterpret_cast<uintptr_t>(exp_avg) |
                reinterpret_cast<uintptr_t>(exp_avg_sq)) & 15) == 0, FMC_ERR_SHAPE, "fmc_adamw_step_f32: buffers must be 16-byte aligned");
  FMC_REQUIRE(step >= 0 && beta1 >= 0.f && beta1 < 1.f && beta2 >= 0.f && beta2 < 1.f, FMC_ERR_ARG,
              "fmc_adamw_step_f32: step must be >= 0 and the betas in [0, 1)");
  FMC_REQUIRE(step >= 1 || state != nullptr, FMC_ERR_ARG, "fmc_adamw_step_f32: step = 0 reads the step count from state[3]");
  FMC_REQUIRE(lr >= 0.f || step == 0, FMC_ERR_ARG, "fmc_adamw_step_f32: lr < 0 (read state[4]) needs step = 0");
  if (n == 0) return FMC_OK;
  const double bc1 = step >= 1 ? 1.0 - pow(static_cast<double>(beta1), step) : 1.0;
  const double bc2 = step >= 1 ? 1.0 - pow(static_cast<double>(beta2), step) : 1.0;
  const long long vecs = (n + 3) / 4;
  launch_k(adamw_step_kernel, dim3(static_cast<unsigned>((vecs + 255) / 256)), dim3(256), 0, stream, param, grad, exp_avg,
           exp_avg_sq, n, lr, beta1, beta2, eps, weight_decay, static_cast<float>(bc1), static_cast<float>(1.0 / sqrt(bc2)),
           state, step == 0 ? 1 : 0);
  return check_launch("adamw_step_kernel");
}
