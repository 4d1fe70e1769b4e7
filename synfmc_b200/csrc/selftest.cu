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

// ---------------------------------------------------------------------------------------------------------
// attention references
// ---------------------------------------------------------------------------------------------------------
__global__ void naive_spatial_attn(const __nv_bfloat16* Q, long long ldq, int q_col0, const __nv_bfloat16* K,
                                   long long ldk, int k_col0, int head_stride, const __nv_bfloat16* V, long long ldv,
                                   int v_col0, float* O, int images, int heads, int d, int nq, int nk, int kv_div,
                                   int kv_stride, float scale) {
  long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  long long total = (long long)images * heads * nq;
  if (idx >= total) return;
  int qi = idx % nq;
  int h = (idx / nq) % heads;
  int img = idx / ((long long)nq * heads);
  const __nv_bfloat16* q = Q + ((long long)img * nq + qi) * ldq + q_col0 + h * head_stride;
  long long kv0 = (long long)(img / kv_div) * kv_stride;
  float mx = -1e30f;
  for (int j = 0; j < nk; ++j) {
    const __nv_bfloat16* k = K + (kv0 + j) * ldk + k_col0 + h * head_stride;
    float s = 0.f;
    for (int c = 0; c < d; ++c) s += __bfloat162float(q[c]) * __bfloat162float(k[c]);
    mx = fmaxf(mx, s * scale);
  }
  float l = 0.f;
  float acc[160];
  for (int c = 0; c < d; ++c) acc[c] = 0.f;
  for (int j = 0; j < nk; ++j) {
    const __nv_bfloat16* k = K + (kv0 + j) * ldk + k_col0 + h * head_stride;
    float s = 0.f;
    for (int c = 0; c < d; ++c) s += __bfloat162float(q[c]) * __bfloat162float(k[c]);
    float pj = expf(s * scale - mx);
    l += pj;
    const __nv_bfloat16* v = V + (kv0 + j) * ldv + v_col0 + h * d;
    for (int c = 0; c < d; ++c) acc[c] += pj * __bfloat162float(v[c]);
  }
  float* o = O + ((long long)img * nq + qi) * (heads * d) + h * d;
  for (int c = 0; c < d; ++c) o[c] = acc[c] / l;
}

__global__ void naive_temporal_attn(const __nv_bfloat16* QKV, long long ld, int q_col0, int k_col0, int v_col0,
                                    int head_stride, float* O, int B, int F, int HW, int heads, int d, float scale) {
  long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  long long total = (long long)B * HW * heads * F;
  if (idx >= total) return;
  int fq = idx % F;
  int h = (idx / F) % heads;
  int hw = (idx / ((long long)F * heads)) % HW;
  int b = idx / ((long long)F * heads * HW);
  auto row = [&](int f) { return ((long long)b * F + f) * HW + hw; };
  const __nv_bfloat16* q = QKV + row(fq) * ld + q_col0 + h * head_stride;
  float s[32];
  float mx = -1e30f;
  for (int j = 0; j < F; ++j) {
    const __nv_bfloat16* k = QKV + row(j) * ld + k_col0 + h * head_stride;
    float a = 0.f;
    for (int c = 0; c < d; ++c) a += __bfloat162float(q[c]) * __bfloat162float(k[c]);
    s[j] = a * scale;
    mx = fmaxf(mx, s[j]);
  }
  float l = 0.f;
  for (int j = 0; j < F; ++j) { s[j] = expf(s[j] - mx); l += s[j]; }
  float* o = O + row(fq) * (heads * d) + h * d;
  for (int c = 0; c < d; ++c) {
    float a = 0.f;
    for (int j = 0; j < F; ++j) a += s[j] * __bfloat162float(QKV[row(j) * ld + v_col0 + h * d + c]);
    o[c] = a / l;
  }
}

static int compare_bf16_vs_f32(const char* tag, const std::vector<__nv_bfloat16>& got, const std::vector<float>& want,
                               double atol, double rtol) {
  double max_err = 0, max_ref = 0, num = 0, den = 0;
  size_t bad = 0;
  for (size_t i = 0; i < want.size(); ++i) {
    double g = __bfloat162float(got[i]), w = want[i];
    double e = fabs(g - w);
    if (!(e <= atol + rtol * fabs(w))) ++bad;
    if (e > max_err || std::isnan(e)) max_err = e;
    if (fabs(w) > max_ref) max_ref = fabs(w);
    num += (g - w) * (g - w);
    den += w * w;
  }
  printf("  %s : max_err=%.4g max|ref|=%.3g rel_l2=%.3g bad=%zu/%zu %s\n", tag, max_err, max_ref, sqrt(num / (den + 1e-30)),
         bad, want.size(), bad == 0 ? "PASS" : "FAIL");
  return bad == 0 ? 0 : 1;
}

// images x nq queries; self (kv_div = 1, kv_stride = nq, nk = nq) or cross (kv_div = f, kv_stride = 80, nk = 77)
static int run_spatial_case(int images, int heads, int d, int nq, int nk, int kv_div, int kv_stride, bool perf) {
  const int hs = (d + 15) / 16 * 16;  // padded Q / K head stride
  const bool self = kv_div == 1 && kv_stride == nq;
  const int C = heads * d;
  const long long q_rows = (long long)images * nq;
  const long long kv_rows = self ? q_rows : (long long)((images + kv_div - 1) / kv_div) * kv_stride;
  // self: one fused buffer [rows, 2*heads*hs + C];  cross: Q buffer [q_rows, heads*hs], KV buffer [kv_rows, heads*hs + C]
  const long long ldq = self ? 2 * heads * hs + C : heads * hs;
  const long long ldk = self ? ldq : heads * hs + C;
  const int q_col0 = 0, k_col0 = self ? heads * hs : 0, v_col0 = self ? 2 * heads * hs : heads * hs;
  std::vector<__nv_bfloat16> hQ((size_t)q_rows * ldq), hKV;
  auto fill = [&](std::vector<__nv_bfloat16>& buf, long long rows, long long ld, int col0, int stride, int width, int valid, float amp) {
    for (long long r = 0; r < rows; ++r)
      for (int h = 0; h < heads; ++h)
        for (int c = 0; c < width; ++c)
          buf[(size_t)(r * ld + col0 + h * stride + c)] = __float2bfloat16(c < valid ? frand() * amp : 0.f);
  };
  for (auto& v : hQ) v = __float2bfloat16(0.f);
  fill(hQ, q_rows, ldq, q_col0, hs, hs, d, 4.0f);
  if (self) {
    fill(hQ, q_rows, ldq, k_col0, hs, hs, d, 4.0f);
    fill(hQ, q_rows, ldq, v_col0, d, d, d, 2.0f);
  } else {
    hKV.assign((size_t)kv_rows * ldk, __float2bfloat16(0.f));
    fill(hKV, kv_rows, ldk, k_col0, hs, hs, d, 4.0f);
    fill(hKV, kv_rows, ldk, v_col0, d, d, d, 2.0f);
    // rows >= nk inside each kv group are padding: keep them zero like the projection of zero-padded text would
    for (long long r = 0; r < kv_rows; ++r)
      if (r % kv_stride >= nk)
        for (long long c = 0; c < ldk; ++c) hKV[(size_t)(r * ldk + c)] = __float2bfloat16(0.f);
  }
  __nv_bfloat16 *dQ, *dKV = nullptr, *dO;
  float* dRef;
  CK(cudaMalloc(&dQ, hQ.size() * 2));
  CK(cudaMemcpy(dQ, hQ.data(), hQ.size() * 2, cudaMemcpyHostToDevice));
  if (!self) {
    CK(cudaMalloc(&dKV, hKV.size() * 2));
    CK(cudaMemcpy(dKV, hKV.data(), hKV.size() * 2, cudaMemcpyHostToDevice));
  }
  const __nv_bfloat16* Kp = self ? dQ : dKV;
  CK(cudaMalloc(&dO, (size_t)q_rows * C * 2));
  CK(cudaMemset(dO, 0xFF, (size_t)q_rows * C * 2));
  CK(cudaMalloc(&dRef, (size_t)q_rows * C * 4));
  const float scale = 1.0f / sqrtf((float)d);
  auto launch = [&]() {
    return fmc_spatial_attn_bf16(dQ, ldq, q_col0, q_rows, Kp, ldk, k_col0, Kp, ldk, v_col0, kv_rows, hs, dO, C, images,
                                 heads, d, nq, nk, kv_div, kv_stride, scale, nullptr);
  };
  int rc = launch();
  if (rc != 0) {
    printf("  spatial launch rc=%d: %s\n", rc, fmc_last_error_string());
    return 1;
  }
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    printf("  spatial kernel failed: %s\n", cudaGetErrorString(e));
    exit(3);
  }
  long long total = q_rows * heads;
  naive_spatial_attn<<<(unsigned)((total + 127) / 128), 128>>>(dQ, ldq, q_col0, Kp, ldk, k_col0, hs, Kp, ldk, v_col0, dRef,
                                                                images, heads, d, nq, nk, kv_div, kv_stride, scale);
  CK(cudaGetLastError());
  CK(cudaDeviceSynchronize());
  std::vector<float> ref((size_t)q_rows * C);
  std::vector<__nv_bfloat16> out((size_t)q_rows * C);
  CK(cudaMemcpy(ref.data(), dRef, ref.size() * 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(out.data(), dO, out.size() * 2, cudaMemcpyDeviceToHost));
  char tag[256];
  snprintf(tag, sizeof(tag), "spatial img=%d heads=%d d=%d nq=%d nk=%d kv_div=%d", images, heads, d, nq, nk, kv_div);
  int fail = compare_bf16_vs_f32(tag, out, ref, 2e-2, 2e-2);
  if (perf && !fail) {
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    for (int i = 0; i < 2; ++i) launch();
    CK(cudaEventRecord(e0));
    const int iters = 10;
    for (int i = 0; i < iters; ++i) launch();
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    ms /= iters;
    printf("    perf: %.3f ms  %.1f TFLOP/s (4*N^2*d per head)\n", ms,
           4.0 * images * heads * (double)nq * nk * d / (ms * 1e-3) / 1e12);
  }
  cudaFree(dQ); cudaFree(dKV); cudaFree(dO); cudaFree(dRef);
  return fail;
}

static int run_temporal_case(int B, int F, int HW, int heads, int d, bool perf) {
  const int hs = (d + 15) / 16 * 16;
  const int C = heads * d;
  const long long rows = (long long)B * F * HW;
  const long long ld = 2 * heads * hs + C;
  const int q_col0 = 0, k_col0 = heads * hs, v_col0 = 2 * heads * hs;
  std::vector<__nv_bfloat16> h((size_t)rows * ld, __float2bfloat16(0.f));
  for (long long r = 0; r < rows; ++r)
    for (int hd = 0; hd < heads; ++hd) {
      for (int c = 0; c < d; ++c) {
        h[(size_t)(r * ld + q_col0 + hd * hs + c)] = __float2bfloat16(frand() * 4.f);
        h[(size_t)(r * ld + k_col0 + hd * hs + c)] = __float2bfloat16(frand() * 4.f);
        h[(size_t)(r * ld + v_col0 + hd * d + c)] = __float2bfloat16(frand() * 2.f);
      }
    }
  __nv_bfloat16 *dX, *dO;
  float* dRef;
  CK(cudaMalloc(&dX, h.size() * 2));
  CK(cudaMemcpy(dX, h.data(), h.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMalloc(&dO, (size_t)rows * C * 2));
  CK(cudaMemset(dO, 0xFF, (size_t)rows * C * 2));
  CK(cudaMalloc(&dRef, (size_t)rows * C * 4));
  const float scale = 1.0f / sqrtf((float)d);
  auto launch = [&]() {
    return fmc_temporal_attn_bf16(dX, ld, q_col0, k_col0, v_col0, hs, dO, C, B, F, HW, heads, d, scale, nullptr);
  };
  int rc = launch();
  if (rc != 0) {
    printf("  temporal launch rc=%d: %s\n", rc, fmc_last_error_string());
    return 1;
  }
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    printf("  temporal kernel failed: %s\n", cudaGetErrorString(e));
    exit(3);
  }
  long long total = (long long)B * HW * heads * F;
  naive_temporal_attn<<<(unsigned)((total + 127) / 128), 128>>>(dX, ld, q_col0, k_col0, v_col0, hs, dRef, B, F, HW, heads, d, scale);
  CK(cudaGetLastError());
  CK(cudaDeviceSynchronize());
  std::vector<float> ref((size_t)rows * C);
  std::vector<__nv_bfloat16> out((size_t)rows * C);
  CK(cudaMemcpy(ref.data(), dRef, ref.size() * 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(out.data(), dO, out.size() * 2, cudaMemcpyDeviceToHost));
  char tag[256];
  snprintf(tag, sizeof(tag), "temporal B=%d F=%d HW=%d heads=%d d=%d", B, F, HW, heads, d);
  int fail = compare_bf16_vs_f32(tag, out, ref, 2e-2, 2e-2);
  if (perf && !fail) {
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    for (int i = 0; i < 2; ++i) launch();
    CK(cudaEventRecord(e0));
    const int iters = 10;
    for (int i = 0; i < iters; ++i) launch();
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    ms /= iters;
    printf("    perf: %.3f ms  %.1f GB/s (q,k,v read + o write)\n", ms, (double)rows * C * 2 * 4 / (ms * 1e-3) / 1e9);
  }
  cudaFree(dX); cudaFree(dO); cudaFree(dRef);
  return fail;
}

int main(int argc, char** argv) {
  bool perf = false;
  const char* which = "all";
  for (int i = 1; i < argc; ++i) {
    if (!strcmp(argv[i], "--perf")) perf = true;
    else which = argv[i];
  }
  const bool all = !strcmp(which, "all");
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  printf("device: %s sm_%d%d, %d SMs, abi %d\n", prop.name, prop.major, prop.minor, prop.multiProcessorCount,
         fmc_abi_version());
  int fails = 0;
  if (all || !strcmp(which, "spatial")) {
    printf("[spatial attention]\n");
    fails += run_spatial_case(2, 8, 40, 256, 256, 1, 256, false);
    fails += run_spatial_case(2, 8, 40, 300, 300, 1, 300, false);
    fails += run_spatial_case(2, 8, 80, 640, 640, 1, 640, false);
    fails += run_spatial_case(3, 8, 160, 160, 160, 1, 160, false);
    fails += run_spatial_case(2, 8, 160, 40, 40, 1, 40, false);
    fails += run_spatial_case(3, 8, 160, 300, 300, 1, 300, false);
    fails += run_spatial_case(4, 8, 40, 200, 77, 2, 80, false);
    fails += run_spatial_case(4, 8, 80, 640, 77, 2, 80, false);
    fails += run_spatial_case(4, 8, 160, 160, 77, 4, 80, false);
    if (perf) {
      fails += run_spatial_case(32, 8, 40, 2560, 2560, 1, 2560, true);
      fails += run_spatial_case(32, 8, 80, 640, 640, 1, 640, true);
      fails += run_spatial_case(32, 8, 160, 160, 160, 1, 160, true);
      fails += run_spatial_case(32, 8, 40, 2560, 77, 16, 80, true);
    }
  }
  if (all || !strcmp(which, "temporal")) {
    printf("[temporal attention]\n");
    fails += run_temporal_case(1, 16, 64, 8, 40, false);
    fails += run_temporal_case(2, 16, 100, 8, 40, false);
    fails += run_temporal_case(2, 16, 50, 8, 80, false);
    fails += run_temporal_case(2, 16, 21, 8, 160, false);
    fails += run_temporal_case(1, 4, 70, 8, 40, false);
    fails += run_temporal_case(1, 8, 33, 8, 80, false);
    fails += run_temporal_case(1, 32, 10, 8, 40, false);
    if (perf) {
      fails += run_temporal_case(2, 16, 2560, 8, 40, true);
      fails += run_temporal_case(2, 16, 640, 8, 80, true);
      fails += run_temporal_case(2, 16, 160, 8, 160, true);
    }
  }
  if (!(all || !strcmp(which, "gemm"))) {
    printf("%s (%d failing cases)\n", fails == 0 ? "SELFTEST PASS" : "SELFTEST FAIL", fails);
    return fails == 0 ? 0 : 1;
  }
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
