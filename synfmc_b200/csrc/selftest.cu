// Standalone GPU self-test of libfmc_b200 kernels (no torch): each kernel is checked against a naive
// CUDA-core implementation on the same device.  Built by `make selftest`; run under gpurun.
// Usage: fmc_selftest [gemm|all] [--perf]
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../include/fmc_b200.h"

#define CK(x)                                                                          \
  do {                                                                                 \
    cudaError_t e_ = (x);                                                              \
    if (e_ != cudaSuccess) {                                                           \
      printf("CUDA error %s at %s:%d: %s\n", #x, __FILE__, __LINE__, cudaGetErrorString(e_)); \
      exit(2);                                                                         \
    }                                                                                  \
  } while (0)

static uint32_t g_seed = 12345u;
static float frand() {
  g_seed = g_seed * 1664525u + 1013904223u;
  return ((g_seed >> 8) & 0xFFFF) / 65536.0f - 0.5f;
}

__global__ void naive_gemm(const __nv_bfloat16* A, long long lda, const __nv_bfloat16* W, long long ldw, float* C,
                           int M, int N, int K) {
  int n = blockIdx.y * blockDim.x + threadIdx.x;
  int m = blockIdx.x;
  if (n >= N || m >= M) return;
  float acc = 0.f;
  for (int k = 0; k < K; ++k) acc += __bfloat162float(A[(long long)m * lda + k]) * __bfloat162float(W[(long long)n * ldw + k]);
  C[(long long)m * N + n] = acc;
}

static float bf16r(float x) { return __bfloat162float(__float2bfloat16(x)); }

struct GemmCase {
  int M, N, K;
  bool bias, residual, rowbias, geglu, f32;
  int tile_n;
};

static int run_gemm_case(const GemmCase& c, bool perf) {
  const int M = c.M, N = c.N, K = c.K;
  const int No = c.geglu ? N / 2 : N;
  std::vector<__nv_bfloat16> hA((size_t)M * K), hW((size_t)N * K), hR((size_t)M * No);
  std::vector<float> hb(N), hrb;
  for (auto& v : hA) v = __float2bfloat16(frand());
  for (auto& v : hW) v = __float2bfloat16(frand() * 0.25f);
  for (auto& v : hR) v = __float2bfloat16(frand());
  for (auto& v : hb) v = frand();
  const int rpg = 37;
  const int groups = (M + rpg - 1) / rpg;
  hrb.resize((size_t)groups * N);
  for (auto& v : hrb) v = frand();

  __nv_bfloat16 *dA, *dW, *dR;
  float *db, *drb, *dRef;
  void* dC;
  CK(cudaMalloc(&dA, hA.size() * 2));
  CK(cudaMalloc(&dW, hW.size() * 2));
  CK(cudaMalloc(&dR, hR.size() * 2));
  CK(cudaMalloc(&db, hb.size() * 4));
  CK(cudaMalloc(&drb, hrb.size() * 4));
  CK(cudaMalloc(&dRef, (size_t)M * N * 4));
  CK(cudaMalloc(&dC, (size_t)M * No * 4));
  CK(cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dW, hW.data(), hW.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dR, hR.data(), hR.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(db, hb.data(), hb.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(drb, hrb.data(), hrb.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemset(dC, 0xFF, (size_t)M * No * 4));

  int flags = (c.geglu ? FMC_GEMM_GEGLU : 0) | (c.f32 ? FMC_GEMM_OUT_F32 : 0);
  int rc = fmc_gemm_bf16(dA, K, dW, K, dC, No, M, N, K, c.bias ? db : nullptr, c.residual ? dR : nullptr, No,
                         c.rowbias ? drb : nullptr, rpg, N, flags, c.tile_n, nullptr);
  if (rc != 0) {
    printf("  gemm launch rc=%d: %s\n", rc, fmc_last_error_string());
    return 1;
  }
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    printf("  gemm kernel failed: %s\n", cudaGetErrorString(e));
    exit(3);  // context is dead after a trap
  }
  naive_gemm<<<dim3(M, (N + 127) / 128), 128>>>(dA, K, dW, K, dRef, M, N, K);
  CK(cudaGetLastError());
  CK(cudaDeviceSynchronize());

  std::vector<float> ref((size_t)M * N);
  CK(cudaMemcpy(ref.data(), dRef, ref.size() * 4, cudaMemcpyDeviceToHost));
  std::vector<float> out((size_t)M * No);
  if (c.f32) {
    CK(cudaMemcpy(out.data(), dC, out.size() * 4, cudaMemcpyDeviceToHost));
  } else {
    std::vector<__nv_bfloat16> ob((size_t)M * No);
    CK(cudaMemcpy(ob.data(), dC, ob.size() * 2, cudaMemcpyDeviceToHost));
    for (size_t i = 0; i < ob.size(); ++i) out[i] = __bfloat162float(ob[i]);
  }
  double max_err = 0, max_ref = 0;
  size_t bad = 0;
  for (int m = 0; m < M; ++m) {
    for (int n = 0; n < No; ++n) {
      double want;
      if (c.geglu) {
        // interleaved blocks of 16: value col = (n/16)*32 + n%16, gate col = value col + 16
        int vc = (n / 16) * 32 + (n % 16);
        double a = ref[(size_t)m * N + vc] + (c.bias ? hb[vc] : 0.f);
        double g = ref[(size_t)m * N + vc + 16] + (c.bias ? hb[vc + 16] : 0.f);
        want = a * 0.5 * g * (1.0 + erf(g / sqrt(2.0)));
      } else {
        want = ref[(size_t)m * N + n] + (c.bias ? hb[n] : 0.f);
        if (c.rowbias) want += hrb[(size_t)(m / rpg) * N + n];
      }
      if (c.residual) want += __bfloat162float(hR[(size_t)m * No + n]);
      double got = out[(size_t)m * No + n];
      double err = fabs(got - want);
      double tol = c.f32 ? 1e-3 + 1e-4 * fabs(want) : 2e-2 + 8e-3 * fabs(want);
      if (!(err <= tol)) ++bad;
      if (err > max_err || std::isnan(err)) max_err = err;
      if (fabs(want) > max_ref) max_ref = fabs(want);
    }
  }
  printf("  gemm M=%d N=%d K=%d bn=%d bias=%d res=%d rowb=%d geglu=%d f32=%d : max_err=%.4g (max|ref|=%.3g) bad=%zu %s\n",
         M, N, K, c.tile_n, c.bias, c.residual, c.rowbias, c.geglu, c.f32, max_err, max_ref, bad,
         bad == 0 ? "PASS" : "FAIL");

  if (perf && bad == 0) {
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    for (int i = 0; i < 3; ++i)
      fmc_gemm_bf16(dA, K, dW, K, dC, No, M, N, K, c.bias ? db : nullptr, c.residual ? dR : nullptr, No, nullptr, 1, 0,
                    flags, c.tile_n, nullptr);
    CK(cudaEventRecord(e0));
    const int iters = 20;
    for (int i = 0; i < iters; ++i)
      fmc_gemm_bf16(dA, K, dW, K, dC, No, M, N, K, c.bias ? db : nullptr, c.residual ? dR : nullptr, No, nullptr, 1, 0,
                    flags, c.tile_n, nullptr);
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    ms /= iters;
    printf("    perf: %.3f ms  %.1f TFLOP/s\n", ms, 2.0 * M * N * K / (ms * 1e-3) / 1e12);
  }
  cudaFree(dA); cudaFree(dW); cudaFree(dR); cudaFree(db); cudaFree(drb); cudaFree(dRef); cudaFree(dC);
  return bad == 0 ? 0 : 1;
}

int main(int argc, char** argv) {
  bool perf = false;
  for (int i = 1; i < argc; ++i)
    if (!strcmp(argv[i], "--perf")) perf = true;
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  printf("device: %s sm_%d%d, %d SMs, abi %d\n", prop.name, prop.major, prop.minor, prop.multiProcessorCount,
         fmc_abi_version());
  int fails = 0;
  printf("[gemm]\n");
  const GemmCase small[] = {
      {128, 128, 64, false, false, false, false, true, 128},
      {128, 128, 128, false, false, false, false, true, 128},
      {128, 64, 64, false, false, false, false, true, 64},
      {128, 160, 320, false, false, false, false, true, 160},
      {128, 256, 256, false, false, false, false, true, 256},
      {256, 320, 320, true, false, false, false, true, 0},
      {200, 320, 320, true, true, false, false, false, 0},
      {77, 640, 768, true, false, false, false, false, 0},
      {2, 1280, 320, true, false, false, false, true, 0},
      {1000, 2560, 320, true, false, false, true, false, 0},
      {1000, 1280, 5120, true, true, true, false, false, 0},
      {5000, 960, 320, false, false, false, false, false, 0},
      {4096, 320, 1280, true, true, false, false, false, 128},
  };
  for (const auto& c : small) fails += run_gemm_case(c, false);
  if (perf) {
    const GemmCase big[] = {
        {81920, 320, 320, true, true, false, false, false, 0},
        {81920, 960, 320, false, false, false, false, false, 0},
        {81920, 2560, 320, true, false, false, true, false, 0},
        {81920, 320, 1280, true, true, false, false, false, 0},
        {20480, 1280, 1280, true, false, false, false, false, 0},
        {20480, 1280, 1280, true, false, false, false, false, 128},
        {20480, 1280, 1280, true, false, false, false, false, 256},
        {8192, 8192, 8192, false, false, false, false, false, 256},
        {8192, 8192, 8192, false, false, false, false, false, 128},
    };
    for (const auto& c : big) fails += run_gemm_case(c, true);
  }
  printf("%s (%d failing cases)\n", fails == 0 ? "SELFTEST PASS" : "SELFTEST FAIL", fails);
  return fails == 0 ? 0 : 1;
}
