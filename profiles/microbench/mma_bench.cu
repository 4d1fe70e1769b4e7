// tcgen05.mma issue / execution timing on one SM, and a layout check of the A-from-TMEM (TS) form.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I../../synfmc_b200/csrc mma_bench.cu -o mma_bench
// Prints cycles per MMA for N in {48, 128, 144, 256}, operands SS / TS, one accumulator chain vs two alternating.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cmath>
#include "ptx.cuh"

using namespace fmc;

__device__ __forceinline__ void umma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}

// mode 0: SS, mode 1: TS.  chains: 1 = every MMA accumulates into the same D, 2 = alternate between two D tiles.
__global__ void __launch_bounds__(128, 1) bench_kernel(int N, int mode, int chains, int reps, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sA = base, sB = base + 16384;
  for (uint32_t off = threadIdx.x * 16; off < 16384 + 32768; off += 128 * 16) st_shared_v4(base + off, 0, 0, 0, 0);
  fence_proxy_async_smem();
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (threadIdx.x < 32) tmem_alloc(&tmem_slot, 512);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = tmem_slot;
  if (threadIdx.x < 32 && elect_one()) {
    const uint32_t idesc = umma_idesc_bf16(128, N);
    const uint64_t da = umma_desc_k_sw128(sA), db = umma_desc_k_sw128(sB);
    // warm-up
    for (int i = 0; i < 8; ++i) umma_bf16_ss(tmem, da, db, idesc, 1);
    umma_commit(&bar);
    mbar_wait(&bar, 0);
    const uint32_t d1 = tmem + (chains == 2 ? 256u : 0u);
    const uint32_t ta = tmem + 448u;
    const long long t0 = clock64();
    if (mode == 0) {
#pragma unroll 1
      for (int i = 0; i < reps; i += 4) {
        umma_bf16_ss(tmem, da, db, idesc, 1);
        umma_bf16_ss(d1, da + 2, db + 2, idesc, 1);
        umma_bf16_ss(tmem, da + 4, db + 4, idesc, 1);
        umma_bf16_ss(d1, da + 6, db + 6, idesc, 1);
      }
    } else if (mode == 2) {
      const uint32_t idesc_mn = umma_idesc_bf16_bmn(128, N);
#pragma unroll 1
      for (int i = 0; i < reps; i += 4) {
        umma_bf16_ts(tmem, ta, umma_desc_mn_sw128(sB, 16384, 1024), idesc_mn, 1);
        umma_bf16_ts(d1, ta + 8, umma_desc_mn_sw128(sB + 2048, 16384, 1024), idesc_mn, 1);
        umma_bf16_ts(tmem, ta + 16, umma_desc_mn_sw128(sB + 4096, 16384, 1024), idesc_mn, 1);
        umma_bf16_ts(d1, ta + 24, umma_desc_mn_sw128(sB + 6144, 16384, 1024), idesc_mn, 1);
      }
    } else {
#pragma unroll 1
      for (int i = 0; i < reps; i += 4) {
        umma_bf16_ts(tmem, ta, db, idesc, 1);
        umma_bf16_ts(d1, ta + 8, db + 2, idesc, 1);
        umma_bf16_ts(tmem, ta + 16, db + 4, idesc, 1);
        umma_bf16_ts(d1, ta + 24, db + 6, idesc, 1);
      }
    }
    const long long t1 = clock64();
    umma_commit(&bar);
    mbar_wait(&bar, 1);
    const long long t2 = clock64();
    out[0] = t1 - t0;
    out[1] = t2 - t0;
  }
  tc_fence_before_sync();
  __syncthreads();
  if (threadIdx.x < 32) { tc_fence_after_sync(); tmem_dealloc(tmem, 512); }
}

// TS layout check: A[128 x 16] lives in TMEM columns [64, 72): lane = row, 32-bit column c holds (k = 2c, 2c + 1).
// B[N=32 x 16] in smem (K-major SW128).  D = A B^T must equal the host product.
__global__ void __launch_bounds__(128, 1) ts_check_kernel(const __nv_bfloat16* A, const __nv_bfloat16* B, float* D) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int t = threadIdx.x;
  for (uint32_t off = t * 16; off < 16384; off += 128 * 16) st_shared_v4(base + off, 0, 0, 0, 0);
  __syncthreads();
  if (t < 32) {  // B row t: 16 bf16 = 2 chunks of 16 B, SW128: chunk j at (j ^ (row & 7))
    const uint4* src = reinterpret_cast<const uint4*>(B + t * 16);
    for (int j = 0; j < 2; ++j) {
      const uint4 v = src[j];
      st_shared_v4(base + t * 128 + ((j ^ (t & 7)) << 4), v.x, v.y, v.z, v.w);
    }
  }
  fence_proxy_async_smem();
  if (t == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (t < 32) tmem_alloc(&tmem_slot, 512);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = tmem_slot;
  const uint32_t lane_addr = tmem + (static_cast<uint32_t>((t >> 5) * 32) << 16);
  {  // each thread stores its row of A: 8 packed words into columns 64..71 (x16 store: upper 8 words are padding)
    uint32_t r[16];
    const uint32_t* src = reinterpret_cast<const uint32_t*>(A + t * 16);
    for (int i = 0; i < 8; ++i) r[i] = src[i];
    for (int i = 8; i < 16; ++i) r[i] = 0;
    tmem_st_x16(lane_addr + 64, r);
    tmem_st_wait();
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  if (t == 0) {
    umma_bf16_ts(tmem, tmem + 64, umma_desc_k_sw128(base), umma_idesc_bf16(128, 32), 0);
    umma_commit(&bar);
  }
  mbar_wait(&bar, 0);
  tc_fence_after_sync();
  uint32_t d[32];
  tmem_ld_x32(lane_addr, d);
  tmem_ld_wait();
  for (int i = 0; i < 32; ++i) D[t * 32 + i] = __uint_as_float(d[i]);
  tc_fence_before_sync();
  __syncthreads();
  if (t < 32) { tc_fence_after_sync(); tmem_dealloc(tmem, 512); }
}

// Mixed operand formats: A = fp16, B = bf16 (both K-major in smem), D = fp32.
__global__ void __launch_bounds__(128, 1) mixed_check_kernel(const __half* A, const __nv_bfloat16* B, float* D) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sA = base, sB = base + 16384;
  const int t = threadIdx.x;
  for (uint32_t off = t * 16; off < 32768; off += 128 * 16) st_shared_v4(base + off, 0, 0, 0, 0);
  __syncthreads();
  {  // A row t: 16 halves = 2 chunks
    const uint4* src = reinterpret_cast<const uint4*>(A + t * 16);
    for (int j = 0; j < 2; ++j) { const uint4 v = src[j]; st_shared_v4(sA + t * 128 + ((j ^ (t & 7)) << 4), v.x, v.y, v.z, v.w); }
  }
  if (t < 32) {
    const uint4* src = reinterpret_cast<const uint4*>(B + t * 16);
    for (int j = 0; j < 2; ++j) { const uint4 v = src[j]; st_shared_v4(sB + t * 128 + ((j ^ (t & 7)) << 4), v.x, v.y, v.z, v.w); }
  }
  fence_proxy_async_smem();
  if (t == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (t < 32) tmem_alloc(&tmem_slot, 512);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = tmem_slot;
  if (t < 32 && elect_one()) {
    // idesc: D = F32 (bit 4), A format = F16 (0 at bits 7-9), B format = BF16 (1 at bits 10-12)
    const uint32_t idesc = (1u << 4) | (0u << 7) | (1u << 10) | (static_cast<uint32_t>(32 >> 3) << 17) | (static_cast<uint32_t>(128 >> 4) << 24);
    umma_bf16_ss(tmem, umma_desc_k_sw128(sA), umma_desc_k_sw128(sB), idesc, 0);
    umma_commit(&bar);
  }
  mbar_wait(&bar, 0);
  tc_fence_after_sync();
  const uint32_t lane_addr = tmem + (static_cast<uint32_t>((t >> 5) * 32) << 16);
  uint32_t d[32];
  tmem_ld_x32(lane_addr, d);
  tmem_ld_wait();
  for (int i = 0; i < 32; ++i) D[t * 32 + i] = __uint_as_float(d[i]);
  tc_fence_before_sync();
  __syncthreads();
  if (t < 32) { tc_fence_after_sync(); tmem_dealloc(tmem, 512); }
}

// ex2.approx.f16x2 availability / accuracy
__global__ void ex2h_kernel(const float* x, float* y, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (2 * i + 1 < n) {
    uint32_t h, r;
    asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(h) : "f"(x[2 * i + 1]), "f"(x[2 * i]));
    asm("ex2.approx.f16x2 %0, %1;" : "=r"(r) : "r"(h));
    const __half2 v = *reinterpret_cast<__half2*>(&r);
    y[2 * i] = __low2float(v);
    y[2 * i + 1] = __high2float(v);
  }
}

int main() {
  long long* out;
  cudaMalloc(&out, 16);
  cudaFuncSetAttribute(bench_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 51200);
  cudaFuncSetAttribute(ts_check_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 20480);
  const int reps = 256;
  for (int mode = 0; mode < 3; ++mode)
    for (int chains = 1; chains <= 2; ++chains)
      for (int N : {48, 64, 128, 256}) {
        if (chains == 2 && N > 192) continue;  // second D tile at column 256, TS A operand at 448..479
        if (mode == 2 && N > 64) continue;      // MN-major B: one 64-element N block
        bench_kernel<<<1, 128, 51200>>>(N, mode, chains, reps, out);
        long long h[2];
        cudaError_t e = cudaMemcpy(h, out, 16, cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) { printf("mode %d N %d: %s\n", mode, N, cudaGetErrorString(e)); return 1; }
        printf("%s chains=%d N=%3d: issue %.1f cyc/MMA, complete %.1f cyc/MMA (ideal %d)\n", mode == 0 ? "SS" : (mode == 1 ? "TS" : "TS/B-MN"), chains, N,
               double(h[0]) / reps, double(h[1]) / reps, N / 2);
      }
  // TS layout check
  std::vector<__nv_bfloat16> hA(128 * 16), hB(32 * 16);
  std::vector<float> fA(128 * 16), fB(32 * 16);
  srand(1);
  for (int i = 0; i < 128 * 16; ++i) { fA[i] = float(rand() % 17 - 8) / 8.f; hA[i] = __float2bfloat16(fA[i]); }
  for (int i = 0; i < 32 * 16; ++i) { fB[i] = float(rand() % 17 - 8) / 8.f; hB[i] = __float2bfloat16(fB[i]); }
  __nv_bfloat16 *dA, *dB;
  float* dD;
  cudaMalloc(&dA, hA.size() * 2); cudaMalloc(&dB, hB.size() * 2); cudaMalloc(&dD, 128 * 32 * 4);
  cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice);
  ts_check_kernel<<<1, 128, 20480>>>(dA, dB, dD);
  std::vector<float> hD(128 * 32);
  cudaError_t e = cudaMemcpy(hD.data(), dD, hD.size() * 4, cudaMemcpyDeviceToHost);
  if (e != cudaSuccess) { printf("ts_check: %s\n", cudaGetErrorString(e)); return 1; }
  double maxerr = 0;
  for (int m = 0; m < 128; ++m)
    for (int n = 0; n < 32; ++n) {
      float ref = 0;
      for (int k = 0; k < 16; ++k) ref += fA[m * 16 + k] * fB[n * 16 + k];
      maxerr = std::max(maxerr, double(fabsf(ref - hD[m * 32 + n])));
    }
  printf("TS layout check (lane = row, column c = k pair (2c, 2c+1)): max |err| = %g  %s\n", maxerr,
         maxerr < 1e-3 ? "OK" : "MISMATCH");
  // (A = fp16 with B = bf16 in one tcgen05.mma kind::f16 was tried here: "illegal instruction" -- the operand
  // formats must match)
  {  // f16x2 exp2
    const int n = 4096;
    std::vector<float> hx(n), hy(n);
    for (int i = 0; i < n; ++i) hx[i] = -20.0f * i / n;
    float *dx, *dy;
    cudaMalloc(&dx, n * 4); cudaMalloc(&dy, n * 4);
    cudaMemcpy(dx, hx.data(), n * 4, cudaMemcpyHostToDevice);
    ex2h_kernel<<<n / 2 / 128, 128>>>(dx, dy, n);
    cudaMemcpy(hy.data(), dy, n * 4, cudaMemcpyDeviceToHost);
    double worst = 0, worst_hi = 0;
    for (int i = 0; i < n; ++i) {
      const double ref = exp2(double(hx[i]));
      const double rel = fabs(hy[i] - ref) / ref;
      if (hx[i] > -13) worst = std::max(worst, rel);
      if (hx[i] > -4) worst_hi = std::max(worst_hi, rel);
    }
    printf("ex2.approx.f16x2: max rel err %.4f (x > -13), %.5f (x > -4)\n", worst, worst_hi);
  }
  return 0;
}
