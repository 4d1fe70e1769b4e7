// Reference-precision mode of the FMC denoising step (BASELINE config 1: "output parity vs reference", 1e-3):
// every activation stays fp32 between kernels, the linears run on the tensor cores as tcgen05.mma.kind::tf32 --
// one pass (operands read as tf32) or the three-pass split  a*w ~= a_hi*w_hi + a_lo*w_hi + a_hi*w_lo  with
// a_hi = tf32(a), a_lo = a - a_hi (fp32-class products, fp32 accumulation in tensor memory) -- attention, norms and
// the glue are fp32 SIMT kernels.  Same arithmetic and reference lines as the bf16 entry points they shadow
// (include/fmc_b200.h cites them per function); nothing here is tuned for speed beyond coalesced, vectorised access.
#include <cfloat>

#include "common.cuh"
#include "ptx.cuh"

namespace fmc {

// =====================================================================================================================
// GEMM  C[M, N] = epilogue(A[M, K] W[N, K]^T), fp32 in / out, tcgen05.mma.kind::tf32 (128 x BN x 8 per instruction)
//   warp 0 TMA producer (SWIZZLE_128B rows of 32 fp32), warp 1 MMA issuer, warp 2 TMEM allocator, warps 4-7 epilogue;
//   persistent CTAs, double-buffered fp32 accumulators in tensor memory.
// nseg = 3: the operands arrive split as [hi | lo] column halves (two tensor maps each); the k-loop runs over the
// segments (A_lo, W_hi), (A_hi, W_lo), (A_hi, W_hi) into ONE accumulator.
// =====================================================================================================================
constexpr int PG_BM = 128;
constexpr int PG_BK = 32;  // fp32 elements per 128-byte swizzle row
constexpr int PG_THREADS = 256;

struct PGemmParams {
  int M, N, K;
  int nseg;
  float* C;
  long long ldc;
  const float* bias;
  const float* residual;
  long long ldr;
  const float* rowbias;
  int rows_per_group;
  long long ldrb;
  int flags;
  int tiles_m, tiles_n;
};

template <int BN>
struct PGemmCfg {
  static constexpr int A_BYTES = PG_BM * PG_BK * 4;
  static constexpr int B_BYTES = BN * PG_BK * 4;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = BN <= 128 ? 5 : 4;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024;
  static constexpr int TMEM_COLS = 2 * BN <= 64 ? 64 : (2 * BN <= 128 ? 128 : (2 * BN <= 256 ? 256 : 512));
  static_assert(BN % 16 == 0 && BN >= 16 && BN <= 256, "UMMA N constraint for M = 128");
};

__host__ __device__ constexpr uint32_t umma_idesc_tf32(int M, int N) {
  return (1u << 4)                                // D format = F32
         | (2u << 7)                              // A format = TF32
         | (2u << 10)                             // B format = TF32
         | (static_cast<uint32_t>(N >> 3) << 17)  // N / 8
         | (static_cast<uint32_t>(M >> 4) << 24); // M / 16
}

__device__ __forceinline__ void umma_tf32_ss(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}

__device__ __forceinline__ float gelu_erf_f32(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }

template <int BN>
__global__ void __launch_bounds__(PG_THREADS, 1)
gemm_tf32_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmA1,
                 const __grid_constant__ CUtensorMap tmB0, const __grid_constant__ CUtensorMap tmB1, PGemmParams p) {
  pdl_launch_dependents();
  using Cfg = PGemmCfg<BN>;
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
  const int kblocks = (p.K + PG_BK - 1) / PG_BK;
  const int kiters = kblocks * p.nseg;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA0);
    tma_prefetch_desc(&tmB0);
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
  if (warp == 2) tmem_alloc(&tmem_base_slot, Cfg::TMEM_COLS);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = tmem_base_slot;
  pdl_wait();

  if (warp == 0) {
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int m_blk = tile / p.tiles_n;
        const int n_blk = tile % p.tiles_n;
        for (int it = 0; it < kiters; ++it) {
          const int seg = it / kblocks;
          const int kb = it - seg * kblocks;
          // three-pass order: (A_lo, W_hi), (A_hi, W_lo), (A_hi, W_hi); one pass: (A, W)
          const CUtensorMap* mA = (p.nseg == 3 && seg == 0) ? &tmA1 : &tmA0;
          const CUtensorMap* mB = (p.nseg == 3 && seg == 1) ? &tmB1 : &tmB0;
          mbar_wait(&empty_bar[stage], phase ^ 1u);
          mbar_arrive_expect_tx(&full_bar[stage], Cfg::STAGE_BYTES);
          const uint32_t sa = smem_base + stage * Cfg::STAGE_BYTES;
          const uint32_t sb = sa + Cfg::A_BYTES;
          tma_load_2d_a(sa, mA, &full_bar[stage], kb * PG_BK, m_blk * PG_BM);
          tma_load_2d_a(sb, mB, &full_bar[stage], kb * PG_BK, n_blk * BN);
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1u;
          }
        }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      constexpr uint32_t idesc = umma_idesc_tf32(PG_BM, BN);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        mbar_wait(&acc_empty_bar[acc], acc_phase ^ 1u);
        tc_fence_after_sync();
        const uint32_t tmem_d = tmem_base + static_cast<uint32_t>(acc * BN);
        for (int it = 0; it < kiters; ++it) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after_sync();
          const uint32_t sa = smem_base + stage * Cfg::STAGE_BYTES;
          const uint32_t sb = sa + Cfg::A_BYTES;
          const uint64_t da = umma_desc_k_sw128(sa);
          const uint64_t db = umma_desc_k_sw128(sb);
#pragma unroll
          for (int k = 0; k < PG_BK / 8; ++k) {
            // +32 bytes per 8-element (tf32) K step inside the 128-byte swizzle span
            umma_tf32_ss(tmem_d, da + static_cast<uint64_t>(2 * k), db + static_cast<uint64_t>(2 * k), idesc,
                         (it > 0 || k > 0) ? 1u : 0u);
          }
          umma_commit(&empty_bar[stage]);
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1u;
          }
        }
        umma_commit(&acc_full_bar[acc]);
        if (++acc == 2) {
          acc = 0;
          acc_phase ^= 1u;
        }
      }
    }
  } else if (warp >= 4) {
    const int q = warp & 3;
    const bool geglu = (p.flags & FMC_GEMM_GEGLU) != 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int m_blk = tile / p.tiles_n;
      const int n_blk = tile % p.tiles_n;
      mbar_wait(&acc_full_bar[acc], acc_phase);
      tc_fence_after_sync();
      const int row = m_blk * PG_BM + q * 32 + lane;
      const bool row_ok = row < p.M;
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(acc * BN);
      const float* rb = nullptr;
      if (p.rowbias != nullptr && row_ok) rb = p.rowbias + static_cast<long long>(row / p.rows_per_group) * p.ldrb;
      if (!geglu) {
#pragma unroll 1
        for (int c = 0; c < BN / 16; ++c) {
          const int col0 = n_blk * BN + c * 16;
          if (col0 >= p.N) break;  // warp-uniform (N % 16 == 0)
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
              const float4* rp = reinterpret_cast<const float4*>(p.residual + static_cast<long long>(row) * p.ldr + col0);
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const float4 u = __ldg(rp + j);
                v[4 * j] += u.x; v[4 * j + 1] += u.y; v[4 * j + 2] += u.z; v[4 * j + 3] += u.w;
              }
            }
            float4* op = reinterpret_cast<float4*>(p.C + static_cast<long long>(row) * p.ldc + col0);
#pragma unroll
            for (int j = 0; j < 4; ++j) op[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
          }
        }
      } else {
        // GEGLU: W rows interleaved in blocks of 16 (value block, gate block); out = value * gelu_erf(gate)
#pragma unroll 1
        for (int c = 0; c < BN / 32; ++c) {
          const int col0 = n_blk * BN + c * 32;
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
            v[j] = a * gelu_erf_f32(g);
          }
          const int ocol0 = col0 >> 1;
          if (row_ok) {
            float4* op = reinterpret_cast<float4*>(p.C + static_cast<long long>(row) * p.ldc + ocol0);
#pragma unroll
            for (int j = 0; j < 4; ++j) op[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
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

template <int BN>
static int launch_gemm_tf32(const CUtensorMap& a0, const CUtensorMap& a1, const CUtensorMap& b0, const CUtensorMap& b1,
                            PGemmParams& p, cudaStream_t stream) {
  using Cfg = PGemmCfg<BN>;
  p.tiles_m = ceil_div(p.M, PG_BM);
  p.tiles_n = ceil_div(p.N, BN);
  static unsigned long long attr_devs = 0;  // per device: the attribute belongs to the (device, function) pair
  if (first_use_on_this_device(&attr_devs)) {
    FMC_CUDA_OK(cudaFuncSetAttribute(gemm_tf32_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
  }
  const int tiles = p.tiles_m * p.tiles_n;
  const int grid = tiles < device_sm_count() ? tiles : device_sm_count();
  FMC_CUDA_OK(launch_k(gemm_tf32_kernel<BN>, dim3(grid), dim3(PG_THREADS), Cfg::SMEM_BYTES, stream, a0, a1, b0, b1, p));
  return check_launch("gemm_tf32_kernel");
}

// out[r, 0:K] = tf32_round_nearest(x[r, :]),  out[r, K:2K] = x - hi   (exact in fp32)
__global__ void __launch_bounds__(256)
split_tf32_kernel(const float* __restrict__ x, long long ldx, float* __restrict__ out, long long ldo, long long rows, int K) {
  pdl_launch_dependents();
  pdl_wait();
  const int nvec = K >> 2;
  const long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (idx >= rows * nvec) return;
  const long long r = idx / nvec;
  const int vi = static_cast<int>(idx % nvec);
  const float4 v = __ldg(reinterpret_cast<const float4*>(x + r * ldx) + vi);
  const float in[4] = {v.x, v.y, v.z, v.w};
  float hi[4], lo[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const uint32_t b = __float_as_uint(in[j]);
    hi[j] = __uint_as_float((b + 0x1000u) & 0xFFFFE000u);  // round half away on the 13 dropped bits
    lo[j] = in[j] - hi[j];
  }
  float* o = out + r * ldo;
  *(reinterpret_cast<float4*>(o) + vi) = make_float4(hi[0], hi[1], hi[2], hi[3]);
  *(reinterpret_cast<float4*>(o + K) + vi) = make_float4(lo[0], lo[1], lo[2], lo[3]);
}

// =====================================================================================================================
// fp32 attention (flash style, SIMT):  O = softmax(Q K^T * scale) V per (image, head).
// A block = 4 warps = 32 queries of one (image, head); key / value tiles of 32 rows in shared memory.  A warp owns 8
// queries: lane j scores key j of the tile against the 8 queries (q broadcast from smem), the online-softmax update
// runs on warp shuffles, then lane c accumulates output channels c, c + 32, ... over the 32 keys (p broadcast).
// Row of element t of sequence i:  (i / inner) * len * inner + (i % inner) + t * inner  -- inner = 1: contiguous
// sequences (spatial tokens of an image); inner = HW: the frame axis of channels-last [B, F, HW, C] (temporal).
// =====================================================================================================================
struct AttnF32Params {
  const float* Q; long long ldq; int q_col0;
  const float* K; long long ldk; int k_col0;
  const float* V; long long ldv; int v_col0;
  float* O; long long ldo;
  int images, heads, nq, nk, kv_div, kv_stride, inner;
  float scale;
};

template <int D>
__global__ void __launch_bounds__(128)
attention_f32_kernel(AttnF32Params p) {
  pdl_launch_dependents();
  pdl_wait();
  constexpr int NCC = (D + 31) / 32;
  constexpr int KS = D + 1;  // padded key row: lane j reads Ks[j][c] conflict-free
  extern __shared__ float smem_f[];
  float* Ks = smem_f;                      // [32][KS]
  float* Vs = Ks + 32 * KS;                // [32][D]
  float* Qs = Vs + 32 * D;                 // [4][D][8]
  float* Ps = Qs + 4 * D * 8;              // [4][32][8]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int img = blockIdx.z, head = blockIdx.y;
  const int q0 = blockIdx.x * 32 + warp * 8;
  const long long q_base = static_cast<long long>(img / p.inner) * p.nq * p.inner + (img % p.inner);
  const int g = img / p.kv_div;
  const long long kv_base = static_cast<long long>(g / p.inner) * p.kv_stride * p.inner + (g % p.inner);

  float* Qw = Qs + warp * D * 8;
  float* Pw = Ps + warp * 32 * 8;
#pragma unroll
  for (int r = 0; r < 8; ++r) {
    const int t = min(q0 + r, p.nq - 1);
    const float* qp = p.Q + (q_base + static_cast<long long>(t) * p.inner) * p.ldq + p.q_col0 + head * D;
    for (int c = lane; c < D; c += 32) Qw[c * 8 + r] = __ldg(qp + c) * p.scale;
  }
  float o[8][NCC];
  float m[8], l[8];
#pragma unroll
  for (int r = 0; r < 8; ++r) {
    m[r] = -INFINITY;
    l[r] = 0.f;
#pragma unroll
    for (int cc = 0; cc < NCC; ++cc) o[r][cc] = 0.f;
  }

  for (int kt = 0; kt < p.nk; kt += 32) {
    __syncthreads();  // previous tile fully consumed (also orders the Qs writes before the first use)
    for (int idx = threadIdx.x; idx < 32 * (D / 4); idx += 128) {
      const int j = idx / (D / 4), c4 = idx % (D / 4);
      float4 kv4 = make_float4(0.f, 0.f, 0.f, 0.f), vv4 = kv4;
      if (kt + j < p.nk) {
        const long long row = kv_base + static_cast<long long>(kt + j) * p.inner;
        kv4 = __ldg(reinterpret_cast<const float4*>(p.K + row * p.ldk + p.k_col0 + head * D) + c4);
        vv4 = __ldg(reinterpret_cast<const float4*>(p.V + row * p.ldv + p.v_col0 + head * D) + c4);
      }
      float* kd = Ks + j * KS + c4 * 4;
      kd[0] = kv4.x; kd[1] = kv4.y; kd[2] = kv4.z; kd[3] = kv4.w;
      *reinterpret_cast<float4*>(Vs + j * D + c4 * 4) = vv4;
    }
    __syncthreads();
    float s[8];
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
    const bool valid = kt + lane < p.nk;
    float pr[8];
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      float sv = valid ? s[r] : -INFINITY;
      float mx = sv;
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, off));
      const float m_new = fmaxf(m[r], mx);
      const float corr = expf(m[r] - m_new);  // first tile: exp(-inf) = 0
      const float pv = valid ? expf(sv - m_new) : 0.f;
      float sum = pv;
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, off);
      l[r] = l[r] * corr + sum;
      m[r] = m_new;
#pragma unroll
      for (int cc = 0; cc < NCC; ++cc) o[r][cc] *= corr;
      pr[r] = pv;
    }
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
        const float v = Vs[j * D + (c < D ? c : 0)];
        o[0][cc] = fmaf(pa.x, v, o[0][cc]); o[1][cc] = fmaf(pa.y, v, o[1][cc]);
        o[2][cc] = fmaf(pa.z, v, o[2][cc]); o[3][cc] = fmaf(pa.w, v, o[3][cc]);
        o[4][cc] = fmaf(pb.x, v, o[4][cc]); o[5][cc] = fmaf(pb.y, v, o[5][cc]);
        o[6][cc] = fmaf(pb.z, v, o[6][cc]); o[7][cc] = fmaf(pb.w, v, o[7][cc]);
      }
    }
    __syncwarp();
  }
#pragma unroll
  for (int r = 0; r < 8; ++r) {
    const int t = q0 + r;
    if (t >= p.nq) continue;
    float* op = p.O + (q_base + static_cast<long long>(t) * p.inner) * p.ldo + head * D;
    const float inv = 1.0f / l[r];
#pragma unroll
    for (int cc = 0; cc < NCC; ++cc) {
      const int c = cc * 32 + lane;
      if (c < D) op[c] = o[r][cc] * inv;
    }
  }
}

template <int D>
static int launch_attention_f32(const AttnF32Params& p, cudaStream_t stream) {
  const int smem = (32 * (D + 1) + 32 * D + 4 * D * 8 + 4 * 32 * 8) * 4;
  static unsigned long long attr_devs = 0;  // per device: the attribute belongs to the (device, function) pair
  if (first_use_on_this_device(&attr_devs)) {
    FMC_CUDA_OK(cudaFuncSetAttribute(attention_f32_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  }
  FMC_CUDA_OK(launch_k(attention_f32_kernel<D>, dim3(ceil_div(p.nq, 32), p.heads, p.images), dim3(128), smem, stream, p));
  return check_launch("attention_f32_kernel");
}

// =====================================================================================================================
// norms, fp32 in / out
// =====================================================================================================================
// LayerNorm, one warp per row, the row in registers; mean then centred variance (torch's formulation)
template <int VPL>  // float4 vectors per lane: C <= VPL * 128
__global__ void __launch_bounds__(256)
layernorm_f32_kernel(const float* __restrict__ x, long long ldx, const float* __restrict__ gamma,
                     const float* __restrict__ beta, float eps, float* __restrict__ out, long long ldo,
                     const float* __restrict__ pe, int F, int HW, const float* __restrict__ add, long long ldadd,
                     float* __restrict__ out2, long long ldo2, long long rows, int C) {
  pdl_launch_dependents();
  pdl_wait();
  const long long row = blockIdx.x * 8ll + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const int nvec = C >> 2;
  float4 v[VPL];
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    const int vi = i * 32 + lane;
    v[i] = vi < nvec ? __ldg(reinterpret_cast<const float4*>(x + row * ldx) + vi) : make_float4(0.f, 0.f, 0.f, 0.f);
    sum += (v[i].x + v[i].y) + (v[i].z + v[i].w);
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, off);
  const float mean = sum / static_cast<float>(C);
  float sq = 0.f;
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    const int vi = i * 32 + lane;
    if (vi < nvec) {
      const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
      sq += (a * a + b * b) + (c * c + d * d);
    }
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, off);
  const float rstd = 1.0f / sqrtf(sq / static_cast<float>(C) + eps);
  const float* pe_row = pe != nullptr ? pe + static_cast<long long>((row / HW) % F) * C : nullptr;
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    const int vi = i * 32 + lane;
    if (vi >= nvec) continue;
    const float4 g4 = __ldg(reinterpret_cast<const float4*>(gamma) + vi);
    const float4 b4 = __ldg(reinterpret_cast<const float4*>(beta) + vi);
    float4 y;
    y.x = (v[i].x - mean) * rstd * g4.x + b4.x;
    y.y = (v[i].y - mean) * rstd * g4.y + b4.y;
    y.z = (v[i].z - mean) * rstd * g4.z + b4.z;
    y.w = (v[i].w - mean) * rstd * g4.w + b4.w;
    if (pe_row != nullptr) {
      const float4 p4 = __ldg(reinterpret_cast<const float4*>(pe_row) + vi);
      y.x += p4.x; y.y += p4.y; y.z += p4.z; y.w += p4.w;
    }
    *(reinterpret_cast<float4*>(out + row * ldo) + vi) = y;
    if (add != nullptr) {
      const float4 a4 = __ldg(reinterpret_cast<const float4*>(add + row * ldadd) + vi);
      *(reinterpret_cast<float4*>(out2 + row * ldo2) + vi) = make_float4(y.x + a4.x, y.y + a4.y, y.z + a4.z, y.w + a4.w);
    }
  }
}

// GroupNorm statistics: one block per (image, group); mean, then centred variance (two reads, fixed reduction order)
__global__ void __launch_bounds__(256)
groupnorm_stats_f32_kernel(const float* __restrict__ x, long long ldx, float eps, float2* __restrict__ stats, int HW, int C,
                           int groups, const float* __restrict__ rowbias, long long ldrb, int rowbias_div) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ float red[8];
  __shared__ float bcast;
  const int img = blockIdx.x / groups, grp = blockIdx.x % groups;
  const int cg = C / groups;
  const float* base = x + static_cast<long long>(img) * HW * ldx + grp * cg;
  const float* rb = rowbias != nullptr ? rowbias + static_cast<long long>(img / rowbias_div) * ldrb + grp * cg : nullptr;
  const int total = HW * cg;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float mean = 0.f;
  for (int pass = 0; pass < 2; ++pass) {
    float acc = 0.f;
    for (int i = threadIdx.x; i < total; i += 256) {
      const int r = i / cg, c = i - r * cg;
      float v = __ldg(base + static_cast<long long>(r) * ldx + c);
      if (rb != nullptr) v += __ldg(rb + c);
      if (pass == 0) acc += v;
      else acc += (v - mean) * (v - mean);
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
    if (lane == 0) red[warp] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
      float t = 0.f;
      for (int w = 0; w < 8; ++w) t += red[w];
      bcast = t / static_cast<float>(total);
    }
    __syncthreads();
    if (pass == 0) mean = bcast;
    else if (threadIdx.x == 0) stats[blockIdx.x] = make_float2(mean, 1.0f / sqrtf(bcast + eps));
    __syncthreads();
  }
}

__global__ void __launch_bounds__(256)
groupnorm_apply_f32_kernel(const float* __restrict__ x, long long ldx, const float* __restrict__ gamma,
                           const float* __restrict__ beta, const float2* __restrict__ stats, float* __restrict__ out,
                           long long ldo, long long rows, int HW, int C, int groups, int silu,
                           const float* __restrict__ rowbias, long long ldrb, int rowbias_div) {
  pdl_launch_dependents();
  pdl_wait();
  const int nvec = C >> 2;
  const long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (idx >= rows * nvec) return;
  const long long r = idx / nvec;
  const int vi = static_cast<int>(idx % nvec);
  const int img = static_cast<int>(r / HW);
  const int cg = C / groups;
  const float4 v4 = __ldg(reinterpret_cast<const float4*>(x + r * ldx) + vi);
  float v[4] = {v4.x, v4.y, v4.z, v4.w};
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int c = vi * 4 + j;
    if (rowbias != nullptr) v[j] += __ldg(rowbias + static_cast<long long>(img / rowbias_div) * ldrb + c);
    const float2 st = __ldg(stats + img * groups + c / cg);
    float y = (v[j] - st.x) * st.y * __ldg(gamma + c) + __ldg(beta + c);
    if (silu) y = y / (1.0f + expf(-y));
    v[j] = y;
  }
  *(reinterpret_cast<float4*>(out + r * ldo) + vi) = make_float4(v[0], v[1], v[2], v[3]);
}

// =====================================================================================================================
// glue, fp32 channels-last
// =====================================================================================================================
// im2col of a 3x3, padding-1 convolution on channels-last [N, H, W, C]: out[(n, oy, ox), (ky, kx, c)]
__global__ void __launch_bounds__(256)
im2col3x3_f32_kernel(const float* __restrict__ x, float* __restrict__ out, int N, int H, int W, int C, int stride, int OH,
                     int OW) {
  pdl_launch_dependents();
  pdl_wait();
  const int nvec = C >> 2;
  const long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  const long long total = static_cast<long long>(N) * OH * OW * 9 * nvec;
  if (idx >= total) return;
  const int vi = static_cast<int>(idx % nvec);
  long long t = idx / nvec;
  const int tap = static_cast<int>(t % 9);
  t /= 9;
  const int ox = static_cast<int>(t % OW);
  t /= OW;
  const int oy = static_cast<int>(t % OH);
  const int n = static_cast<int>(t / OH);
  const int iy = oy * stride + tap / 3 - 1, ix = ox * stride + tap % 3 - 1;
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  if (iy >= 0 && iy < H && ix >= 0 && ix < W)
    v = __ldg(reinterpret_cast<const float4*>(x + ((static_cast<long long>(n) * H + iy) * W + ix) * C) + vi);
  *(reinterpret_cast<float4*>(out) + idx) = v;
}

__global__ void __launch_bounds__(256)
add_f32_kernel(const float* __restrict__ a, long long lda, const float* __restrict__ b, long long ldb,
               const float* __restrict__ rowbias, int rows_per_group, long long ldrb, float* __restrict__ out,
               long long ldo, long long rows, int C, int relu) {
  pdl_launch_dependents();
  pdl_wait();
  const int nvec = C >> 2;
  const long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (idx >= rows * nvec) return;
  const long long r = idx / nvec;
  const int vi = static_cast<int>(idx % nvec);
  float4 v = __ldg(reinterpret_cast<const float4*>(a + r * lda) + vi);
  if (b != nullptr) {
    const float4 w = __ldg(reinterpret_cast<const float4*>(b + r * ldb) + vi);
    v.x += w.x; v.y += w.y; v.z += w.z; v.w += w.w;
  }
  if (rowbias != nullptr) {
    const float4 w = __ldg(reinterpret_cast<const float4*>(rowbias + (r / rows_per_group) * ldrb) + vi);
    v.x += w.x; v.y += w.y; v.z += w.z; v.w += w.w;
  }
  if (relu) {
    v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f);
  }
  *(reinterpret_cast<float4*>(out + r * ldo) + vi) = v;
}

__global__ void __launch_bounds__(256)
resize_nearest_f32_kernel(const float* __restrict__ x, float* __restrict__ out, int N, int h, int w, int oh, int ow, int C,
                          float sh, float sw) {
  pdl_launch_dependents();
  pdl_wait();
  const int nvec = C >> 2;
  const long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  const long long total = static_cast<long long>(N) * oh * ow * nvec;
  if (idx >= total) return;
  const int vi = static_cast<int>(idx % nvec);
  long long t = idx / nvec;
  const int ox = static_cast<int>(t % ow);
  t /= ow;
  const int oy = static_cast<int>(t % oh);
  const int n = static_cast<int>(t / oh);
  const int iy = min(static_cast<int>(floorf(oy * sh)), h - 1);
  const int ix = min(static_cast<int>(floorf(ox * sw)), w - 1);
  *(reinterpret_cast<float4*>(out) + idx) =
      __ldg(reinterpret_cast<const float4*>(x + ((static_cast<long long>(n) * h + iy) * w + ix) * C) + vi);
}

__global__ void __launch_bounds__(256)
avgpool2_f32_kernel(const float* __restrict__ x, float* __restrict__ out, int N, int h, int w, int C) {
  pdl_launch_dependents();
  pdl_wait();
  const int oh = h >> 1, ow = w >> 1, nvec = C >> 2;
  const long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  const long long total = static_cast<long long>(N) * oh * ow * nvec;
  if (idx >= total) return;
  const int vi = static_cast<int>(idx % nvec);
  long long t = idx / nvec;
  const int ox = static_cast<int>(t % ow);
  t /= ow;
  const int oy = static_cast<int>(t % oh);
  const int n = static_cast<int>(t / oh);
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
  for (int dy = 0; dy < 2; ++dy)
#pragma unroll
    for (int dx = 0; dx < 2; ++dx) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(
                                 x + ((static_cast<long long>(n) * h + 2 * oy + dy) * w + 2 * ox + dx) * C) + vi);
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
  *(reinterpret_cast<float4*>(out) + idx) = make_float4(acc.x * 0.25f, acc.y * 0.25f, acc.z * 0.25f, acc.w * 0.25f);
}

__global__ void __launch_bounds__(256)
copy2d_f32_kernel(const float* __restrict__ src, long long lds, float* __restrict__ dst, long long ldd, long long rows,
                  int cols) {
  pdl_launch_dependents();
  pdl_wait();
  const int nvec = cols >> 2;
  const long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (idx >= rows * nvec) return;
  const long long r = idx / nvec;
  const int vi = static_cast<int>(idx % nvec);
  *(reinterpret_cast<float4*>(dst + r * ldd) + vi) = __ldg(reinterpret_cast<const float4*>(src + r * lds) + vi);
}

__global__ void __launch_bounds__(256)
ncfhw_to_cl_f32_kernel(const float* __restrict__ x, float* __restrict__ out, int B, int C, int F, long long HW, int Cpad) {
  pdl_launch_dependents();
  pdl_wait();
  const long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  const long long total = static_cast<long long>(B) * F * HW * Cpad;
  if (idx >= total) return;
  const int c = static_cast<int>(idx % Cpad);
  long long t = idx / Cpad;
  const long long hw = t % HW;
  t /= HW;
  const int f = static_cast<int>(t % F);
  const int b = static_cast<int>(t / F);
  out[idx] = c < C ? __ldg(x + ((static_cast<long long>(b) * C + c) * F + f) * HW + hw) : 0.f;
}

__global__ void __launch_bounds__(256)
cl_to_ncfhw_f32_kernel(const float* __restrict__ x, long long ldc, float* __restrict__ out, int B, int C, int F,
                       long long HW) {
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
  out[idx] = x[((static_cast<long long>(b) * F + f) * HW + hw) * ldc + c];
}

__global__ void __launch_bounds__(256)
silu_f32_kernel(const float* __restrict__ x, float* __restrict__ out, long long n) {
  pdl_launch_dependents();
  pdl_wait();
  const long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (idx >= n) return;
  const float v = x[idx];
  out[idx] = v / (1.0f + expf(-v));
}

__global__ void timestep_embedding_f32_kernel(const float* __restrict__ t, float* __restrict__ out, int B, int dim) {
  pdl_launch_dependents();
  pdl_wait();
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int half = dim / 2;
  if (idx >= B * half) return;
  const int b = idx / half, i = idx % half;
  const float e = expf(-logf(10000.0f) * static_cast<float>(i) / static_cast<float>(half));
  const float a = t[b] * e;
  out[b * dim + i] = cosf(a);
  out[b * dim + half + i] = sinf(a);
}

__global__ void __launch_bounds__(256)
mask_modulate_f32_kernel(const float* __restrict__ x, const float* __restrict__ mask, const int* __restrict__ ry,
                         const int* __restrict__ rx, float* __restrict__ out, int N, int h, int w, int C, int H, int W) {
  pdl_launch_dependents();
  pdl_wait();
  const int nvec = C >> 2;
  const long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  const long long total = static_cast<long long>(N) * h * w * nvec;
  if (idx >= total) return;
  long long t = idx / nvec;
  const int xx = static_cast<int>(t % w);
  t /= w;
  const int yy = static_cast<int>(t % h);
  const int n = static_cast<int>(t / h);
  const float m = __ldg(mask + (static_cast<long long>(n) * H + __ldg(ry + yy)) * W + __ldg(rx + xx));
  const float4 v = __ldg(reinterpret_cast<const float4*>(x) + idx);
  *(reinterpret_cast<float4*>(out) + idx) = make_float4(v.x * m, v.y * m, v.z * m, v.w * m);
}

static inline unsigned pblocks(long long n) { return static_cast<unsigned>((n + 255) / 256); }

}  // namespace fmc

using namespace fmc;

extern "C" int fmc_gemm_tf32(const float* A, long long lda, const float* W, long long ldw, float* C, long long ldc,
                             int M, int N, int K, const float* bias, const float* residual, long long ldr,
                             const float* rowbias, int rows_per_group, long long ldrb, int flags, int split,
                             void* stream) {
  FMC_REQUIRE(A && W && C, FMC_ERR_ARG, "fmc_gemm_tf32: null operand");
  FMC_REQUIRE(split == 1 || split == 3, FMC_ERR_ARG, "fmc_gemm_tf32: split must be 1 or 3");
  FMC_REQUIRE((flags & ~FMC_GEMM_GEGLU) == 0, FMC_ERR_ARG, "fmc_gemm_tf32: only FMC_GEMM_GEGLU is a valid flag");
  const bool geglu = (flags & FMC_GEMM_GEGLU) != 0;
  FMC_REQUIRE(M > 0 && N > 0 && K > 0 && K % 4 == 0 && N % (geglu ? 32 : 16) == 0, FMC_ERR_SHAPE,
              "fmc_gemm_tf32: M=%d N=%d K=%d (K %% 4, N %% %d required)", M, N, K, geglu ? 32 : 16);
  FMC_REQUIRE(lda % 4 == 0 && ldw % 4 == 0 && ldc % 4 == 0 && (residual == nullptr || ldr % 4 == 0), FMC_ERR_SHAPE,
              "fmc_gemm_tf32: row strides must be multiples of 4 floats");
  FMC_REQUIRE(!(geglu && (residual != nullptr || rowbias != nullptr)), FMC_ERR_ARG,
              "fmc_gemm_tf32: GEGLU takes no residual / row bias");
  FMC_REQUIRE(rowbias == nullptr || rows_per_group > 0, FMC_ERR_ARG, "fmc_gemm_tf32: rows_per_group must be positive");
  const int BN = N <= 32 ? 32 : (N <= 64 ? 64 : ((N % 256 == 0 && M >= 2048) ? 256 : 128));
  CUtensorMap a0, a1, b0, b1;
  const uint64_t dA[2] = {static_cast<uint64_t>(K), static_cast<uint64_t>(M)};
  const uint64_t sA[1] = {static_cast<uint64_t>(lda) * 4};
  const uint32_t boxA[2] = {PG_BK, PG_BM};
  const uint64_t dB[2] = {static_cast<uint64_t>(K), static_cast<uint64_t>(N)};
  const uint64_t sB[1] = {static_cast<uint64_t>(ldw) * 4};
  const uint32_t boxB[2] = {PG_BK, static_cast<uint32_t>(BN)};
  int rc = make_tmap_f32(&a0, A, 2, dA, sA, boxA, 128);
  if (rc != FMC_OK) return rc;
  rc = make_tmap_f32(&b0, W, 2, dB, sB, boxB, 128);
  if (rc != FMC_OK) return rc;
  a1 = a0;
  b1 = b0;
  if (split == 3) {
    rc = make_tmap_f32(&a1, A + K, 2, dA, sA, boxA, 128);
    if (rc != FMC_OK) return rc;
    rc = make_tmap_f32(&b1, W + K, 2, dB, sB, boxB, 128);
    if (rc != FMC_OK) return rc;
  }
  PGemmParams p;
  p.M = M; p.N = N; p.K = K; p.nseg = split;
  p.C = C; p.ldc = ldc; p.bias = bias; p.residual = residual; p.ldr = ldr;
  p.rowbias = rowbias; p.rows_per_group = rows_per_group > 0 ? rows_per_group : 1; p.ldrb = ldrb;
  p.flags = flags; p.tiles_m = p.tiles_n = 0;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  switch (BN) {
    case 32: return launch_gemm_tf32<32>(a0, a1, b0, b1, p, s);
    case 64: return launch_gemm_tf32<64>(a0, a1, b0, b1, p, s);
    case 256: return launch_gemm_tf32<256>(a0, a1, b0, b1, p, s);
    default: return launch_gemm_tf32<128>(a0, a1, b0, b1, p, s);
  }
}

extern "C" int fmc_split_tf32(const float* x, long long ldx, float* out, long long ldo, long long rows, int K,
                              void* stream) {
  FMC_REQUIRE(x && out, FMC_ERR_ARG, "fmc_split_tf32: null operand");
  FMC_REQUIRE(K % 4 == 0 && ldx % 4 == 0 && ldo % 4 == 0 && ldo >= 2 * K, FMC_ERR_SHAPE,
              "fmc_split_tf32: K and strides must be multiples of 4, ldo >= 2 K");
  if (rows == 0) return FMC_OK;
  launch_k(split_tf32_kernel, dim3(pblocks(rows * (K / 4))), dim3(256), 0, static_cast<cudaStream_t>(stream), x, ldx, out,
           ldo, rows, K);
  return check_launch("split_tf32_kernel");
}

extern "C" int fmc_attention_f32(const float* Q, long long ldq, int q_col0, const float* K, long long ldk, int k_col0,
                                 const float* V, long long ldv, int v_col0, float* O, long long ldo, int images,
                                 int heads, int head_dim, int nq, int nk, int kv_div, int kv_stride, int inner,
                                 float scale, void* stream) {
  FMC_REQUIRE(Q && K && V && O, FMC_ERR_ARG, "fmc_attention_f32: null operand");
  FMC_REQUIRE(images > 0 && heads > 0 && nq > 0 && nk > 0 && kv_div > 0 && inner > 0, FMC_ERR_ARG,
              "fmc_attention_f32: sizes must be positive");
  FMC_REQUIRE(inner == 1 || kv_div == 1, FMC_ERR_ARG, "fmc_attention_f32: strided sequences take their own keys");
  FMC_REQUIRE(images <= 65535 && heads <= 65535, FMC_ERR_SHAPE, "fmc_attention_f32: grid limit (images=%d)", images);
  FMC_REQUIRE(ldk % 4 == 0 && ldv % 4 == 0 && k_col0 % 4 == 0 && v_col0 % 4 == 0, FMC_ERR_SHAPE,
              "fmc_attention_f32: K / V rows must be 16-byte aligned");
  AttnF32Params p;
  p.Q = Q; p.ldq = ldq; p.q_col0 = q_col0; p.K = K; p.ldk = ldk; p.k_col0 = k_col0; p.V = V; p.ldv = ldv;
  p.v_col0 = v_col0; p.O = O; p.ldo = ldo; p.images = images; p.heads = heads; p.nq = nq; p.nk = nk;
  p.kv_div = kv_div; p.kv_stride = kv_stride; p.inner = inner; p.scale = scale;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  switch (head_dim) {
    case 40: return launch_attention_f32<40>(p, s);
    case 80: return launch_attention_f32<80>(p, s);
    case 160: return launch_attention_f32<160>(p, s);
    default: break;
  }
  set_error("fmc_attention_f32: head_dim %d not in {40, 80, 160}", head_dim);
  return FMC_ERR_SHAPE;
}

extern "C" int fmc_layernorm_f32(const float* x, long long ldx, const float* gamma, const float* beta, float eps,
                                 float* out, long long ldo, const float* pe, int F, int HW, const float* add,
                                 long long ldadd, float* out2, long long ldo2, long long rows, int C, void* stream) {
  FMC_REQUIRE(x && gamma && beta && out, FMC_ERR_ARG, "fmc_layernorm_f32: null operand");
  FMC_REQUIRE(C % 4 == 0 && C <= 2560 && ldx % 4 == 0 && ldo % 4 == 0, FMC_ERR_SHAPE,
              "fmc_layernorm_f32: C=%d must be a multiple of 4, at most 2560", C);
  FMC_REQUIRE((add == nullptr) == (out2 == nullptr), FMC_ERR_ARG, "fmc_layernorm_f32: add and out2 go together");
  FMC_REQUIRE(pe == nullptr || (F > 0 && HW > 0), FMC_ERR_ARG, "fmc_layernorm_f32: pe needs F and HW");
  if (rows == 0) return FMC_OK;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const dim3 grid(static_cast<unsigned>((rows + 7) / 8));
  if (C <= 384) {
    launch_k(layernorm_f32_kernel<3>, grid, dim3(256), 0, s, x, ldx, gamma, beta, eps, out, ldo, pe, F > 0 ? F : 1,
             HW > 0 ? HW : 1, add, ldadd, out2, ldo2, rows, C);
  } else if (C <= 1280) {
    launch_k(layernorm_f32_kernel<10>, grid, dim3(256), 0, s, x, ldx, gamma, beta, eps, out, ldo, pe, F > 0 ? F : 1,
             HW > 0 ? HW : 1, add, ldadd, out2, ldo2, rows, C);
  } else {
    launch_k(layernorm_f32_kernel<20>, grid, dim3(256), 0, s, x, ldx, gamma, beta, eps, out, ldo, pe, F > 0 ? F : 1,
             HW > 0 ? HW : 1, add, ldadd, out2, ldo2, rows, C);
  }
  return check_launch("layernorm_f32_kernel");
}

extern "C" int fmc_groupnorm_f32(const float* x, long long ldx, const float* gamma, const float* beta, float eps,
                                 float* out, long long ldo, float* stats_ws, int images, int HW, int C, int groups,
                                 int silu, const float* rowbias, long long ldrb, int rowbias_div, void* stream) {
  FMC_REQUIRE(x && gamma && beta && out && stats_ws, FMC_ERR_ARG, "fmc_groupnorm_f32: null operand");
  FMC_REQUIRE(groups > 0 && C % groups == 0 && C % 4 == 0 && ldx % 4 == 0 && ldo % 4 == 0, FMC_ERR_SHAPE,
              "fmc_groupnorm_f32: C=%d groups=%d", C, groups);
  FMC_REQUIRE(rowbias == nullptr || rowbias_div > 0, FMC_ERR_ARG, "fmc_groupnorm_f32: rowbias_div must be positive");
  if (images == 0 || HW == 0) return FMC_OK;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  float2* stats = reinterpret_cast<float2*>(stats_ws);
  const int div = rowbias_div > 0 ? rowbias_div : 1;
  launch_k(groupnorm_stats_f32_kernel, dim3(images * groups), dim3(256), 0, s, x, ldx, eps, stats, HW, C, groups, rowbias,
           ldrb, div);
  int rc = check_launch("groupnorm_stats_f32_kernel");
  if (rc != FMC_OK) return rc;
  const long long rows = static_cast<long long>(images) * HW;
  launch_k(groupnorm_apply_f32_kernel, dim3(pblocks(rows * (C / 4))), dim3(256), 0, s, x, ldx, gamma, beta,
           static_cast<const float2*>(stats), out, ldo, rows, HW, C, groups, silu, rowbias, ldrb, div);
  return check_launch("groupnorm_apply_f32_kernel");
}

extern "C" int fmc_im2col3x3_f32(const float* x, float* out, int N, int H, int W, int C, int stride, void* stream) {
  FMC_REQUIRE(x && out, FMC_ERR_ARG, "fmc_im2col3x3_f32: null operand");
  FMC_REQUIRE(C % 4 == 0 && (stride == 1 || stride == 2), FMC_ERR_SHAPE, "fmc_im2col3x3_f32: C=%d stride=%d", C, stride);
  const int OH = (H - 1) / stride + 1, OW = (W - 1) / stride + 1;  // kernel 3, padding 1
  const long long total = static_cast<long long>(N) * OH * OW * 9 * (C / 4);
  if (total == 0) return FMC_OK;
  launch_k(im2col3x3_f32_kernel, dim3(pblocks(total)), dim3(256), 0, static_cast<cudaStream_t>(stream), x, out, N, H, W, C,
           stride, OH, OW);
  return check_launch("im2col3x3_f32_kernel");
}

extern "C" int fmc_add_f32(const float* a, long long lda, const float* b, long long ldb, const float* rowbias,
                           int rows_per_group, long long ldrb, float* out, long long ldo, long long rows, int C, int relu,
                           void* stream) {
  FMC_REQUIRE(a && out, FMC_ERR_ARG, "fmc_add_f32: null operand");
  FMC_REQUIRE(C % 4 == 0 && lda % 4 == 0 && ldo % 4 == 0 && (b == nullptr || ldb % 4 == 0) &&
                  (rowbias == nullptr || ldrb % 4 == 0), FMC_ERR_SHAPE, "fmc_add_f32: C and strides must be multiples of 4");
  FMC_REQUIRE(rowbias == nullptr || rows_per_group > 0, FMC_ERR_ARG, "fmc_add_f32: rows_per_group must be positive");
  if (rows == 0) return FMC_OK;
  launch_k(add_f32_kernel, dim3(pblocks(rows * (C / 4))), dim3(256), 0, static_cast<cudaStream_t>(stream), a, lda, b, ldb,
           rowbias, rows_per_group > 0 ? rows_per_group : 1, ldrb, out, ldo, rows, C, relu);
  return check_launch("add_f32_kernel");
}

extern "C" int fmc_resize_nearest_f32(const float* x, float* out, int N, int h, int w, int oh, int ow, int C,
                                      void* stream) {
  FMC_REQUIRE(x && out, FMC_ERR_ARG, "fmc_resize_nearest_f32: null operand");
  FMC_REQUIRE(C % 4 == 0, FMC_ERR_SHAPE, "fmc_resize_nearest_f32: C must be a multiple of 4");
  const long long total = static_cast<long long>(N) * oh * ow * (C / 4);
  if (total == 0) return FMC_OK;
  launch_k(resize_nearest_f32_kernel, dim3(pblocks(total)), dim3(256), 0, static_cast<cudaStream_t>(stream), x, out, N, h, w,
           oh, ow, C, static_cast<float>(h) / oh, static_cast<float>(w) / ow);
  return check_launch("resize_nearest_f32_kernel");
}

extern "C" int fmc_avgpool2_f32(const float* x, float* out, int N, int h, int w, int C, void* stream) {
  FMC_REQUIRE(x && out, FMC_ERR_ARG, "fmc_avgpool2_f32: null operand");
  FMC_REQUIRE(C % 4 == 0, FMC_ERR_SHAPE, "fmc_avgpool2_f32: C must be a multiple of 4");
  const long long total = static_cast<long long>(N) * (h / 2) * (w / 2) * (C / 4);
  if (total == 0) return FMC_OK;
  launch_k(avgpool2_f32_kernel, dim3(pblocks(total)), dim3(256), 0, static_cast<cudaStream_t>(stream), x, out, N, h, w, C);
  return check_launch("avgpool2_f32_kernel");
}

extern "C" int fmc_copy2d_f32(const float* src, long long lds, float* dst, long long ldd, long long rows, int cols,
                              void* stream) {
  FMC_REQUIRE(src && dst, FMC_ERR_ARG, "fmc_copy2d_f32: null operand");
  FMC_REQUIRE(cols % 4 == 0 && lds % 4 == 0 && ldd % 4 == 0, FMC_ERR_SHAPE, "fmc_copy2d_f32: cols / strides must be multiples of 4");
  if (rows == 0 || cols == 0) return FMC_OK;
  launch_k(copy2d_f32_kernel, dim3(pblocks(rows * (cols / 4))), dim3(256), 0, static_cast<cudaStream_t>(stream), src, lds,
           dst, ldd, rows, cols);
  return check_launch("copy2d_f32_kernel");
}

extern "C" int fmc_ncfhw_f32_to_cl_f32(const float* x, float* out, int B, int C, int F, long long HW, int Cpad,
                                       void* stream) {
  FMC_REQUIRE(x && out && Cpad >= C, FMC_ERR_ARG, "fmc_ncfhw_f32_to_cl_f32: bad arguments");
  const long long total = static_cast<long long>(B) * F * HW * Cpad;
  if (total == 0) return FMC_OK;
  launch_k(ncfhw_to_cl_f32_kernel, dim3(pblocks(total)), dim3(256), 0, static_cast<cudaStream_t>(stream), x, out, B, C, F,
           HW, Cpad);
  return check_launch("ncfhw_to_cl_f32_kernel");
}

extern "C" int fmc_cl_f32_to_ncfhw_f32(const float* x, long long ldc, float* out, int B, int C, int F, long long HW,
                                       void* stream) {
  FMC_REQUIRE(x && out && ldc >= C, FMC_ERR_ARG, "fmc_cl_f32_to_ncfhw_f32: bad arguments");
  const long long total = static_cast<long long>(B) * C * F * HW;
  if (total == 0) return FMC_OK;
  launch_k(cl_to_ncfhw_f32_kernel, dim3(pblocks(total)), dim3(256), 0, static_cast<cudaStream_t>(stream), x, ldc, out, B, C,
           F, HW);
  return check_launch("cl_to_ncfhw_f32_kernel");
}

extern "C" int fmc_silu_f32(const float* x, float* out, long long n, void* stream) {
  FMC_REQUIRE(x && out, FMC_ERR_ARG, "fmc_silu_f32: null operand");
  if (n == 0) return FMC_OK;
  launch_k(silu_f32_kernel, dim3(pblocks(n)), dim3(256), 0, static_cast<cudaStream_t>(stream), x, out, n);
  return check_launch("silu_f32_kernel");
}

extern "C" int fmc_timestep_embedding_f32(const float* t, float* out, int B, int dim, void* stream) {
  FMC_REQUIRE(t && out && dim % 2 == 0, FMC_ERR_ARG, "fmc_timestep_embedding_f32: bad arguments");
  if (B == 0) return FMC_OK;
  launch_k(timestep_embedding_f32_kernel, dim3((B * dim / 2 + 127) / 128), dim3(128), 0, static_cast<cudaStream_t>(stream),
           t, out, B, dim);
  return check_launch("timestep_embedding_f32_kernel");
}

extern "C" int fmc_mask_modulate_f32(const float* x, const float* mask, const int* row_index, const int* col_index,
                                     float* out, int N, int h, int w, int C, int H, int W, void* stream) {
  FMC_REQUIRE(x && mask && row_index && col_index && out, FMC_ERR_ARG, "fmc_mask_modulate_f32: null operand");
  FMC_REQUIRE(C % 4 == 0, FMC_ERR_SHAPE, "fmc_mask_modulate_f32: C must be a multiple of 4");
  const long long total = static_cast<long long>(N) * h * w * (C / 4);
  if (total == 0) return FMC_OK;
  launch_k(mask_modulate_f32_kernel, dim3(pblocks(total)), dim3(256), 0, static_cast<cudaStream_t>(stream), x, mask,
           row_index, col_index, out, N, h, w, C, H, W);
  return check_launch("mask_modulate_f32_kernel");
}
