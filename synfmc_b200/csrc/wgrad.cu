// Weight gradient dW[M, N] (fp32) = dY^T X = sum_t dY[t, m] X[t, n] on tcgen05, straight from the row-major activation
// gradients dY [T, M] and inputs X [T, N] (bf16): BOTH operands are fed MN-major -- a TMA box of 64 tokens x 64 channels
// (SWIZZLE_128B) is exactly the MN-major operand tile of a K = 64 step (the layout the attention kernels use for V) -- so
// the two transposed copies the first version made for a K-major GEMM disappear; and the token axis is SPLIT across
// CTAs (a 320 x 320 gradient over 40 960 tokens is only nine 128 x 128 tiles), each split writing an fp32 partial that a
// second kernel folds in a fixed order (deterministic, no atomics).
//   warp 0  TMA producer (4-stage ring of [A 64 x 128 | B 64 x 128] token slabs)     warp 1  tcgen05.mma issuer
//   warp 2  TMEM allocator (128 fp32 columns)                                         warps 4-7  epilogue (TMEM -> global)
// Replaces the wgrad half of `loss.backward()` for the trainable linears (qkv_merge of the CameraAdapter, the
// CameraEncoder's / ObjectEncoder's linears): train_cam_ctrl.py:648, train_cam_obj_ctrl.py:857.
#include <cuda_bf16.h>

#include "common.cuh"
#include "ptx.cuh"

namespace fmc {

constexpr int WG_BM = 128, WG_BN = 128, WG_BK = 64, WG_STAGES = 4, WG_THREADS = 256;
constexpr int WG_CHUNK = WG_BK * 128;            // [64 tokens x 64 channels] SWIZZLE_128B box
constexpr int WG_STAGE = 4 * WG_CHUNK;           // A: two 64-channel chunks, B: two
constexpr int WG_SMEM = WG_STAGES * WG_STAGE + 1024;

// kind::f16 instruction descriptor, bf16 x bf16 -> fp32, A and B both MN-major
__host__ __device__ constexpr uint32_t umma_idesc_bf16_amn_bmn(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | (static_cast<uint32_t>(N >> 3) << 17) |
         (static_cast<uint32_t>(M >> 4) << 24);
}

struct WgParams {
  int T, M, N, steps_per_split;  // k-steps (of 64 tokens) per split (= blockIdx.z)
  int partial;                   // 1: out = partials [splits, M, N]; 0: out = dW [M, ldo] (single split, no accumulate)
  float* out;
  long long ldo;
};

__global__ void __launch_bounds__(WG_THREADS, 1)
wgrad_tn_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, WgParams p) {
  pdl_launch_dependents();
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t full[WG_STAGES], empty[WG_STAGES], done;
  __shared__ uint32_t tmem_base_slot;
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.x * WG_BM, n0 = blockIdx.y * WG_BN, split = blockIdx.z;
  const int total_steps = (p.T + WG_BK - 1) / WG_BK;
  const int step0 = split * p.steps_per_split;
  const int nsteps = min(p.steps_per_split, total_steps - step0);  // >= 1 by construction of the grid

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < WG_STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    mbar_init(&done, 1);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(&tmem_base_slot, 128);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = tmem_base_slot;
  pdl_wait();

  if (warp == 0) {
    if (elect_one()) {
      for (int i = 0; i < nsteps; ++i) {
        const int st = i % WG_STAGES;
        mbar_wait(&empty[st], ((i / WG_STAGES) & 1u) ^ 1u);
        mbar_arrive_expect_tx(&full[st], WG_STAGE);
        const uint32_t sA = smem_base + st * WG_STAGE, sB = sA + 2 * WG_CHUNK;
        const int t0 = (step0 + i) * WG_BK;
        tma_load_2d_a(sA, &tmA, &full[st], m0, t0);
        tma_load_2d_a(sA + WG_CHUNK, &tmA, &full[st], m0 + 64, t0);
        tma_load_2d_a(sB, &tmB, &full[st], n0, t0);
        tma_load_2d_a(sB + WG_CHUNK, &tmB, &full[st], n0 + 64, t0);
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      constexpr uint32_t idesc = umma_idesc_bf16_amn_bmn(WG_BM, WG_BN);
      for (int i = 0; i < nsteps; ++i) {
        const int st = i % WG_STAGES;
        mbar_wait(&full[st], (i / WG_STAGES) & 1u);
        tc_fence_after_sync();
        const uint32_t sA = smem_base + st * WG_STAGE, sB = sA + 2 * WG_CHUNK;
#pragma unroll
        for (int k = 0; k < WG_BK / 16; ++k) {
          // 16 tokens per MMA: rows k*16 .. of both slabs; the two 64-channel chunks of an operand are WG_CHUNK apart (LBO)
          const uint64_t da = umma_desc_mn_sw128(sA + k * (16 * 128), WG_CHUNK, 1024);
          const uint64_t db = umma_desc_mn_sw128(sB + k * (16 * 128), WG_CHUNK, 1024);
          umma_bf16_ss(tmem_base, da, db, idesc, (i > 0 || k > 0) ? 1u : 0u);
        }
        umma_commit(&empty[st]);
      }
      umma_commit(&done);
    }
  } else if (warp >= 4) {
    const int q = warp & 3;
    const int r = q * 32 + lane;  // row of the tile = channel m0 + r of dY
    mbar_wait(&done, 0);
    tc_fence_after_sync();
    const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    const int m = m0 + r;
    float* dst = p.partial ? p.out + (static_cast<long long>(split) * p.M + m) * p.N
                           : p.out + static_cast<long long>(m) * p.ldo;
#pragma unroll 1
    for (int c = 0; c < WG_BN / 32; ++c) {
      uint32_t v[32];
      tmem_ld_x32(lane_addr + c * 32, v);
      tmem_ld_wait();
      if (m < p.M) {
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          const int n = n0 + c * 32 + j;
          if (n + 4 <= p.N) {
            *reinterpret_cast<float4*>(dst + n) = make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]),
                                                               __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
          } else {
            for (int e = 0; e < 4; ++e)
              if (n + e < p.N) dst[n + e] = __uint_as_float(v[j + e]);
          }
        }
      }
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after_sync();
    tmem_dealloc(tmem_base, 128);
  }
}

// out[m, n] (+)= sum_s part[s, m, n], splits folded in index order
__global__ void __launch_bounds__(256)
wgrad_fold_kernel(const float* __restrict__ part, float* __restrict__ out, long long ldo, int M, int N, int splits,
                  int accumulate) {
  pdl_launch_dependents();
  pdl_wait();
  const long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  const int nv = N >> 2;
  if (idx >= static_cast<long long>(M) * nv) return;
  const int m = static_cast<int>(idx / nv), v = static_cast<int>(idx % nv);
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int s = 0; s < splits; ++s) {
    const float4 x = __ldg(reinterpret_cast<const float4*>(part + (static_cast<long long>(s) * M + m) * N) + v);
    acc.x += x.x; acc.y += x.y; acc.z += x.z; acc.w += x.w;
  }
  float4* o = reinterpret_cast<float4*>(out + static_cast<long long>(m) * ldo) + v;
  if (accumulate) {
    const float4 y = *o;
    acc.x += y.x; acc.y += y.y; acc.z += y.z; acc.w += y.w;
  }
  *o = acc;
}

static int wgrad_splits(long long T, int M, int N) {
  const int tiles = ceil_div(M, WG_BM) * ceil_div(N, WG_BN);
  const int steps = static_cast<int>((T + WG_BK - 1) / WG_BK);
  int splits = ceil_div(2 * device_sm_count(), tiles);       // about two CTAs per SM
  const int max_splits = steps >= 8 ? steps / 8 : 1;         // at least 8 k-steps (512 tokens) per split
  splits = splits < 1 ? 1 : (splits > max_splits ? max_splits : splits);
  const int per = ceil_div(steps, splits);
  return ceil_div(steps, per);                               // no empty split
}

}  // namespace fmc

using namespace fmc;

extern "C" long long fmc_wgrad_workspace_floats(long long T, int M, int N) {
  return static_cast<long long>(wgrad_splits(T, M, N)) * M * N;
}

extern "C" int fmc_wgrad_bf16(const void* dY, long long lddy, const void* X, long long ldx, float* dW, long long lddw,
                              float* workspace, long long T, int M, int N, int accumulate, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  FMC_REQUIRE(dY && X && dW, FMC_ERR_ARG, "fmc_wgrad_bf16: null operand");
  FMC_REQUIRE(T > 0 && M > 0 && N > 0 && T < 0x7fffffffll, FMC_ERR_SHAPE, "fmc_wgrad_bf16: empty problem");
  FMC_REQUIRE(M % 8 == 0 && N % 8 == 0 && lddy % 8 == 0 && ldx % 8 == 0 && lddw % 4 == 0 &&
                  (reinterpret_cast<uintptr_t>(dY) & 15) == 0 && (reinterpret_cast<uintptr_t>(X) & 15) == 0 &&
                  (reinterpret_cast<uintptr_t>(dW) & 15) == 0, FMC_ERR_SHAPE,
              "fmc_wgrad_bf16: M=%d, N=%d and the row strides must be multiples of 8 (16-byte rows)", M, N);
  const int splits = wgrad_splits(T, M, N);
  const bool direct = splits == 1 && !accumulate;
  FMC_REQUIRE(direct || workspace != nullptr, FMC_ERR_ARG,
              "fmc_wgrad_bf16: %d token split(s)%s need the workspace (fmc_wgrad_workspace_floats)", splits,
              accumulate ? " with accumulation" : "");
  CUtensorMap tmA, tmB;
  const uint32_t box[2] = {64, WG_BK};
  {
    const uint64_t dims[2] = {static_cast<uint64_t>(M), static_cast<uint64_t>(T)};
    const uint64_t strides[1] = {static_cast<uint64_t>(lddy) * 2};
    int rc = make_tmap_bf16(&tmA, dY, 2, dims, strides, box, true);
    if (rc != FMC_OK) return rc;
  }
  {
    const uint64_t dims[2] = {static_cast<uint64_t>(N), static_cast<uint64_t>(T)};
    const uint64_t strides[1] = {static_cast<uint64_t>(ldx) * 2};
    int rc = make_tmap_bf16(&tmB, X, 2, dims, strides, box, true);
    if (rc != FMC_OK) return rc;
  }
  const int steps = static_cast<int>((T + WG_BK - 1) / WG_BK);
  WgParams p{};
  p.T = static_cast<int>(T); p.M = M; p.N = N; p.steps_per_split = ceil_div(steps, splits);
  p.partial = direct ? 0 : 1;
  p.out = direct ? dW : workspace;
  p.ldo = lddw;
  static unsigned long long devs = 0;
  if (first_use_on_this_device(&devs))
    FMC_CUDA_OK(cudaFuncSetAttribute(wgrad_tn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, WG_SMEM));
  launch_k(wgrad_tn_kernel, dim3(ceil_div(M, WG_BM), ceil_div(N, WG_BN), splits), dim3(WG_THREADS), WG_SMEM, stream, tmA, tmB, p);
  int rc = check_launch("wgrad_tn_kernel");
  if (rc != FMC_OK || direct) return rc;
  const long long vecs = static_cast<long long>(M) * (N / 4);
  launch_k(wgrad_fold_kernel, dim3(static_cast<unsigned>((vecs + 255) / 256)), dim3(256), 0, stream,
           static_cast<const float*>(workspace), dW, lddw, M, N, splits, accumulate);
  return check_launch("wgrad_fold_kernel");
}
