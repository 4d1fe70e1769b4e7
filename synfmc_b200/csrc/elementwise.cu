// HBM-bound glue kernels of the denoising step on channels-last bf16 activations: vectorised (16-byte) loads and
// stores, one pass over the data each.  Reference lines are cited per entry point in include/fmc_b200.h.
#include <cuda_bf16.h>

#include "common.cuh"
#include "ptx.cuh"

namespace fmc {

__device__ __forceinline__ void unpack8(const uint4& u, float (&v)[8]) {
  const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    v[2 * j] = bf16_lo(w[j]);
    v[2 * j + 1] = bf16_hi(w[j]);
  }
}
__device__ __forceinline__ uint4 pack8(const float (&v)[8]) {
  return make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]));
}

// out[r, c] = a[r, c] (+ b[r, c]) (+ rowbias[r / rows_per_group, c]), optional ReLU.  Strided rows, C % 8 == 0.
__global__ void __launch_bounds__(256)
add_kernel(const __nv_bfloat16* __restrict__ a, long long lda, const __nv_bfloat16* __restrict__ b, long long ldb,
           const float* __restrict__ rowbias, int rows_per_group, long long ldrb, __nv_bfloat16* __restrict__ out,
           long long ldo, long long rows, int C, int relu) {
  pdl_launch_dependents();
  pdl_wait();
  const int nvec = C >> 3;
  const long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (idx >= rows * nvec) return;
  const long long r = idx / nvec;
  const int vi = static_cast<int>(idx % nvec);
  float v[8];
  unpack8(__ldg(reinterpret_cast<const uint4*>(a + r * lda) + vi), v);
  if (b != nullptr) {
    float w[8];
    unpack8(__ldg(reinterpret_cast<const uint4*>(b + r * ldb) + vi), w);
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] += w[j];
  }
  if (rowbias != nullptr) {
    const float* rb = rowbias + (r / rows_per_group) * ldrb + vi * 8;
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] += __ldg(rb + j);
  }
  if (relu) {
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = fmaxf(v[j], 0.f);
  }
  *(reinterpret_cast<uint4*>(out + r * ldo) + vi) = pack8(v);
}

// Nearest-neighbour resize of [N, h, w, C] to [N, oh, ow, C] (torch 'nearest': src = floor(dst * in / out)).
__global__ void __launch_bounds__(256)
resize_nearest_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ out, int N, int h, int w, int oh,
                      int ow, int C, float sh, float sw) {
  pdl_launch_dependents();
  pdl_wait();
  const int nvec = C >> 3;
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
  const uint4 u = __ldg(reinterpret_cast<const uint4*>(x + ((static_cast<long long>(n) * h + iy) * w + ix) * C) + vi);
  *(reinterpret_cast<uint4*>(out + ((static_cast<long long>(n) * oh + oy) * ow + ox) * C) + vi) = u;
}

// 2x2 average pooling (AvgPool2d(2), floor) of [N, h, w, C].
__global__ void __launch_bounds__(256)
avgpool2_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ out, int N, int h, int w, int C) {
  pdl_launch_dependents();
  pdl_wait();
  const int oh = h >> 1, ow = w >> 1, nvec = C >> 3;
  const long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  const long long total = static_cast<long long>(N) * oh * ow * nvec;
  if (idx >= total) return;
  const int vi = static_cast<int>(idx % nvec);
  long long t = idx / nvec;
  const int ox = static_cast<int>(t % ow);
  t /= ow;
  const int oy = static_cast<int>(t % oh);
  const int n = static_cast<int>(t / oh);
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int dy = 0; dy < 2; ++dy)
#pragma unroll
    for (int dx = 0; dx < 2; ++dx) {
      float v[8];
      unpack8(__ldg(reinterpret_cast<const uint4*>(
                        x + ((static_cast<long long>(n) * h + 2 * oy + dy) * w + 2 * ox + dx) * C) + vi), v);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] += v[j];
    }
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] *= 0.25f;
  *(reinterpret_cast<uint4*>(out + ((static_cast<long long>(n) * oh + oy) * ow + ox) * C) + vi) = pack8(acc);
}

// Strided 2-D copy (channel concat of skip connections): dst[r, 0:cols] = src[r, 0:cols].
__global__ void __launch_bounds__(256)
copy2d_kernel(const __nv_bfloat16* __restrict__ src, long long lds, __nv_bfloat16* __restrict__ dst, long long ldd,
              long long rows, int cols) {
  pdl_launch_dependents();
  pdl_wait();
  const int nvec = cols >> 3;
  const long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (idx >= rows * nvec) return;
  const long long r = idx / nvec;
  const int vi = static_cast<int>(idx % nvec);
  *(reinterpret_cast<uint4*>(dst + r * ldd) + vi) = __ldg(reinterpret_cast<const uint4*>(src + r * lds) + vi);
}

// [B, C, F, H, W] fp32 (reference layout) -> [B, F, H, W, Cpad] bf16 (channels-last, zero padded channels).
__global__ void __launch_bounds__(256)
ncfhw_to_cl_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ out, int B, int C, int F, long long HW,
                   int Cpad) {
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
  const float v = c < C ? __ldg(x + ((static_cast<long long>(b) * C + c) * F + f) * HW + hw) : 0.f;
  out[idx] = __float2bfloat16(v);
}

// [B, F, H, W, ldc] bf16 -> [B, C, F, H, W] fp32.
__global__ void __launch_bounds__(256)
cl_to_ncfhw_kernel(const __nv_bfloat16* __restrict__ x, long long ldc, float* __restrict__ out, int B, int C, int F,
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
  out[idx] = __bfloat162float(x[((static_cast<long long>(b) * F + f) * HW + hw) * ldc + c]);
}

// fp32 -> bf16 with optional SiLU (time-embedding activations feeding the projection GEMMs).
__global__ void __launch_bounds__(256)
cast_act_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ out, long long n, int silu) {
  pdl_launch_dependents();
  pdl_wait();
  const long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (idx >= n) return;
  float v = x[idx];
  if (silu) v = v / (1.0f + __expf(-v));
  out[idx] = __float2bfloat16(v);
}

// diffusers Timesteps(dim, flip_sin_to_cos=True, freq_shift=0): out[b] = [cos(t e_i) | sin(t e_i)], e_i = 1e4^(-i/half).
__global__ void timestep_embedding_kernel(const float* __restrict__ t, __nv_bfloat16* __restrict__ out, int B, int dim) {
  pdl_launch_dependents();
  pdl_wait();
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int half = dim / 2;
  if (idx >= B * half) return;
  const int b = idx / half, i = idx % half;
  const float e = expf(-logf(10000.0f) * static_cast<float>(i) / static_cast<float>(half));
  const float a = t[b] * e;
  out[b * dim + i] = __float2bfloat16(cosf(a));
  out[b * dim + half + i] = __float2bfloat16(sinf(a));
}

// ---------------------------------------------------------------------------------------------------------------
// Pluecker rays (a9): K[b*f, 4] = (fx, fy, cx, cy), c2w[b*f, 3, 4].  For pixel centre (i + .5, j + .5):
//   dir = normalize((i - cx) / fx, (j - cy) / fy, 1);  d = dir * R^T;  o = t;  out = (o x d, d)
// plain: out[b*f, H, W, 6] fp32.  unshuffled: out[b*f, H/8, W/8, 6*64] bf16 with channel comp*64 + dy*8 + dx
// (the PixelUnshuffle(8) of CameraPoseEncoder.forward, fmc/models/pose_adaptor.py:227-228, fused away).
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void plucker_pixel(const float* __restrict__ K, const float* __restrict__ M, int px, int py,
                                              float (&out)[6]) {
  const float fx = K[0], fy = K[1], cx = K[2], cy = K[3];
  const float x = ((static_cast<float>(px) + 0.5f) - cx) / fx;
  const float y = ((static_cast<float>(py) + 0.5f) - cy) / fy;
  const float z = 1.0f;
  const float n = sqrtf(x * x + y * y + z * z);
  const float dx = x / n, dy = y / n, dz = z / n;
  // rays_d = dir @ R^T  ->  d_k = sum_m dir_m R[k][m]
  const float d0 = dx * M[0] + dy * M[1] + dz * M[2];
  const float d1 = dx * M[4] + dy * M[5] + dz * M[6];
  const float d2 = dx * M[8] + dy * M[9] + dz * M[10];
  const float o0 = M[3], o1 = M[7], o2 = M[11];
  out[0] = o1 * d2 - o2 * d1;
  out[1] = o2 * d0 - o0 * d2;
  out[2] = o0 * d1 - o1 * d0;
  out[3] = d0;
  out[4] = d1;
  out[5] = d2;
}

__global__ void __launch_bounds__(256)
plucker_plain_kernel(const float* __restrict__ K, const float* __restrict__ c2w, float* __restrict__ out, int BF, int H,
                     int W) {
  pdl_launch_dependents();
  pdl_wait();
  const long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  const long long total = static_cast<long long>(BF) * H * W;
  if (idx >= total) return;
  const int px = static_cast<int>(idx % W);
  const int py = static_cast<int>((idx / W) % H);
  const int n = static_cast<int>(idx / (static_cast<long long>(W) * H));
  float v[6];
  plucker_pixel(K + n * 4, c2w + n * 12, px, py, v);
  float* o = out + idx * 6;
#pragma unroll
  for (int j = 0; j < 6; ++j) o[j] = v[j];
}

// one thread per (frame, cell y, cell x, dy): 8 pixels of one image row -> for each of the 6 components one 16-byte store
// (channels comp*64 + dy*8 + 0..7).  Write-only and HBM bound once the per-pixel arithmetic is out of the way: everything
// that does not depend on the pixel column is hoisted (1/fx, 1/fy, y, y^2 + 1, y R[k][1] + R[k][2]), the normalisation is
// one rsqrt.approx and the direction d_k = r (x R[k][0] + c_k) -- ~20 instructions per pixel instead of ~150 with
// IEEE division and sqrt (the first version was ISSUE bound: 18.4 us = 1.7 TB/s on 31.5 MB, profiles/r02_membound.md).
// The result is rounded to bf16 (2^-9), far above the 2-ulp fp32 approximations; the fp32 entry point
// fmc_plucker_f32 keeps the reference's exact operation order.
__global__ void __launch_bounds__(256)
plucker_unshuffle_kernel(const float* __restrict__ K, const float* __restrict__ c2w, __nv_bfloat16* __restrict__ out,
                         int BF, int H, int W) {
  pdl_launch_dependents();
  pdl_wait();
  const int h8 = H >> 3, w8 = W >> 3;
  const long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  const long long total = static_cast<long long>(BF) * h8 * w8 * 8;
  if (idx >= total) return;
  const int dy = static_cast<int>(idx & 7);
  long long t = idx >> 3;
  const int cx = static_cast<int>(t % w8);
  t /= w8;
  const int cy = static_cast<int>(t % h8);
  const int n = static_cast<int>(t / h8);
  const float4 k4 = __ldg(reinterpret_cast<const float4*>(K) + n);          // fx, fy, cx, cy
  const float4* Mp = reinterpret_cast<const float4*>(c2w) + n * 3;          // rows of [R | t]
  const float4 m0 = __ldg(Mp), m1 = __ldg(Mp + 1), m2 = __ldg(Mp + 2);
  const float inv_fx = 1.0f / k4.x, inv_fy = 1.0f / k4.y;
  const float y = ((static_cast<float>(cy * 8 + dy) + 0.5f) - k4.w) * inv_fy;
  const float yy1 = fmaf(y, y, 1.0f);
  const float c0 = fmaf(y, m0.y, m0.z), c1 = fmaf(y, m1.y, m1.z), c2 = fmaf(y, m2.y, m2.z);
  const float o0 = m0.w, o1 = m1.w, o2 = m2.w;
  const float x0 = ((static_cast<float>(cx * 8) + 0.5f) - k4.z) * inv_fx;
  float v[6][8];
#pragma unroll
  for (int dx = 0; dx < 8; ++dx) {
    const float x = fmaf(static_cast<float>(dx), inv_fx, x0);
    const float r = rsqrtf(fmaf(x, x, yy1));
    const float d0 = r * fmaf(x, m0.x, c0), d1 = r * fmaf(x, m1.x, c1), d2 = r * fmaf(x, m2.x, c2);
    v[0][dx] = fmaf(o1, d2, -o2 * d1);
    v[1][dx] = fmaf(o2, d0, -o0 * d2);
    v[2][dx] = fmaf(o0, d1, -o1 * d0);
    v[3][dx] = d0;
    v[4][dx] = d1;
    v[5][dx] = d2;
  }
  __nv_bfloat16* cell = out + ((static_cast<long long>(n) * h8 + cy) * w8 + cx) * 384 + dy * 8;
#pragma unroll
  for (int comp = 0; comp < 6; ++comp) *reinterpret_cast<uint4*>(cell + comp * 64) = pack8(v[comp]);
}

// ---------------------------------------------------------------------------------------------------------------
// ObjectEncoder input (a10), fmc/util.py:147-203: per pixel the LAST object with mask > 0 wins;
//   traj = (info * m) * m   (12 channels),   maskf = m * m ... exactly: features = cat(info*m, m) * m.
// plain:      feat[bf, 13, H, W] fp32 + mask[bf, H, W] fp32 (reference layout, bit-exact scatter)
// unshuffled: feat[bf, H/8, W/8, 13*64] bf16 (channel c*64 + dy*8 + dx) + mask[bf, H, W] fp32
// info[bf, n_obj, 12] fp32, masks[bf, n_obj, H, W] fp32.
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void traj_pixel(const float* __restrict__ info, const float* __restrict__ masks, int n_obj,
                                           long long hw_stride, long long pix, float (&feat)[13], float& mval) {
  int sel = -1;
  float m = 0.f;
  for (int o = 0; o < n_obj; ++o) {
    const float mo = __ldg(masks + o * hw_stride + pix);
    if (mo > 0.f) {
      sel = o;
      m = mo;
    }
  }
  mval = m;
  if (sel < 0) {
#pragma unroll
    for (int c = 0; c < 13; ++c) feat[c] = 0.f;
    return;
  }
#pragma unroll
  for (int c = 0; c < 12; ++c) {
    const float masked = __fmul_rn(__ldg(info + sel * 12 + c), m);  // expanded_obj_info * obj_mask   (util.py:176)
    feat[c] = __fmul_rn(masked, m);                                 // features * mask_features       (util.py:200)
  }
  feat[12] = __fmul_rn(m, m);
}

__global__ void __launch_bounds__(256)
traj_plain_kernel(const float* __restrict__ info, const float* __restrict__ masks, float* __restrict__ feat,
                  float* __restrict__ mask_out, int BF, int n_obj, int H, int W) {
  pdl_launch_dependents();
  pdl_wait();
  const long long HW = static_cast<long long>(H) * W;
  const long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (idx >= BF * HW) return;
  const long long pix = idx % HW;
  const int n = static_cast<int>(idx / HW);
  float f[13], m;
  traj_pixel(info + static_cast<long long>(n) * n_obj * 12, masks + static_cast<long long>(n) * n_obj * HW, n_obj, HW, pix, f, m);
#pragma unroll
  for (int c = 0; c < 13; ++c) feat[(static_cast<long long>(n) * 13 + c) * HW + pix] = f[c];
  mask_out[idx] = m;
}

// one thread per (frame, cell y, cell x, dy) = 8 pixels of one image row: the masks of every object are read as two
// 16-byte loads, the winning (object, mask value) per pixel stays in registers, and each of the 13 channels leaves as
// one 16-byte store (channels c*64 + dy*8 + 0..7); the combined mask as two 16-byte stores.  Arithmetic as traj_pixel.
__global__ void __launch_bounds__(256)
traj_unshuffle_kernel(const float* __restrict__ info, const float* __restrict__ masks, __nv_bfloat16* __restrict__ feat,
                      float* __restrict__ mask_out, int BF, int n_obj, int H, int W) {
  pdl_launch_dependents();
  pdl_wait();
  const int h8 = H >> 3, w8 = W >> 3;
  const long long HW = static_cast<long long>(H) * W;
  const long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  const long long total = static_cast<long long>(BF) * h8 * w8 * 8;
  if (idx >= total) return;
  const int dy = static_cast<int>(idx & 7);
  long long t = idx >> 3;
  const int cx = static_cast<int>(t % w8);
  t /= w8;
  const int cy = static_cast<int>(t % h8);
  const int n = static_cast<int>(t / h8);
  const long long pix0 = static_cast<long long>(cy * 8 + dy) * W + cx * 8;  // W % 8 == 0: 32-byte aligned
  const float* mrow = masks + static_cast<long long>(n) * n_obj * HW + pix0;
  int sel[8];
  float m[8];
#pragma unroll
  for (int dx = 0; dx < 8; ++dx) {
    sel[dx] = -1;
    m[dx] = 0.f;
  }
  for (int o = 0; o < n_obj; ++o) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(mrow + o * HW));
    const float4 b = __ldg(reinterpret_cast<const float4*>(mrow + o * HW) + 1);
    const float mo[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
    for (int dx = 0; dx < 8; ++dx) {
      if (mo[dx] > 0.f) {  // the LAST object with mask > 0 wins (fmc/util.py:178-182)
        sel[dx] = o;
        m[dx] = mo[dx];
      }
    }
  }
  float4* mo4 = reinterpret_cast<float4*>(mask_out + static_cast<long long>(n) * HW + pix0);
  mo4[0] = make_float4(m[0], m[1], m[2], m[3]);
  mo4[1] = make_float4(m[4], m[5], m[6], m[7]);
  const float* inf = info + static_cast<long long>(n) * n_obj * 12;
  __nv_bfloat16* cell = feat + ((static_cast<long long>(n) * h8 + cy) * w8 + cx) * 832 + dy * 8;
#pragma unroll
  for (int c = 0; c < 12; ++c) {
    float r[8];
#pragma unroll
    for (int dx = 0; dx < 8; ++dx) {
      const float iv = sel[dx] >= 0 ? __ldg(inf + sel[dx] * 12 + c) : 0.f;
      r[dx] = __fmul_rn(__fmul_rn(iv, m[dx]), m[dx]);  // (expanded_obj_info * obj_mask) * mask_features (util.py:176,200)
    }
    *reinterpret_cast<uint4*>(cell + c * 64) = pack8(r);
  }
  float r[8];
#pragma unroll
  for (int dx = 0; dx < 8; ++dx) r[dx] = __fmul_rn(m[dx], m[dx]);
  *reinterpret_cast<uint4*>(cell + 12 * 64) = pack8(r);
}

// ---------------------------------------------------------------------------------------------------------------
// Gaussian sphere masks on the device (fmc/data/dataset.py:5350-5403, `use_sphere_mask`): the dataset
// turns each object's segmentation mask into its minimum enclosing circle (centre, radius; cv2 on the host) and then into
//     gaussian = exp(-0.5 (dist / sigma)^2), sigma = radius / 2;  gaussian /= gaussian.max();  gaussian *= disc
// with the disc rasterised by cv2.circle(int(centre), int(radius), filled).  Everything after the circle is a closed
// form of (cx, cy, r): generating it here removes the [n_obj, H, W] mask tensors from host memory, the H2D copy
// (10 MB per object and clip at 320x512x16f) and -- in the fused form below -- the mask reads of the scatter kernel.
// Disc test: (x - int(cx))^2 + (y - int(cy))^2 <= int(r)^2 (cv2's midpoint rasterisation may differ by single boundary
// pixels; cv2 is not available offline to pin that).  `gaussian.max()` over the image is the value at the pixel nearest
// to the centre.  r <= 0 marks an absent object (empty segmentation mask: the reference keeps the all-zero mask).
// ---------------------------------------------------------------------------------------------------------------
struct SphereObj {
  float cx, cy, inv_s2, norm;  // inv_s2 = -0.5 / sigma^2, norm = 1 / gaussian.max()
  int icx, icy, r2;            // r2 < 0: absent
};
__device__ __forceinline__ SphereObj sphere_obj(const float* __restrict__ c, int H, int W) {
  SphereObj o;
  const float cx = __ldg(c), cy = __ldg(c + 1), r = __ldg(c + 2);
  o.cx = cx; o.cy = cy;
  if (!(r > 0.f)) {
    o.r2 = -1; o.icx = o.icy = 0; o.inv_s2 = 0.f; o.norm = 0.f;
    return o;
  }
  const float sigma = r * 0.5f;
  o.inv_s2 = -0.5f / (sigma * sigma);
  o.icx = static_cast<int>(cx); o.icy = static_cast<int>(cy);
  const int ir = static_cast<int>(r);
  o.r2 = ir * ir;
  const float nx = fminf(fmaxf(rintf(cx), 0.f), static_cast<float>(W - 1));
  const float ny = fminf(fmaxf(rintf(cy), 0.f), static_cast<float>(H - 1));
  const float dm2 = (nx - cx) * (nx - cx) + (ny - cy) * (ny - cy);
  o.norm = 1.0f / expf(dm2 * o.inv_s2);
  return o;
}
__device__ __forceinline__ float sphere_value(const SphereObj& o, int px, int py) {
  const int dxi = px - o.icx, dyi = py - o.icy;
  if (o.r2 < 0 || dxi * dxi + dyi * dyi > o.r2) return 0.f;
  const float dx = static_cast<float>(px) - o.cx, dy = static_cast<float>(py) - o.cy;
  return expf((dx * dx + dy * dy) * o.inv_s2) * o.norm;
}

// masks[bf, o, y, x] from circles[bf, o, 3] = (cx, cy, r)
__global__ void __launch_bounds__(256)
sphere_mask_kernel(const float* __restrict__ circles, float* __restrict__ masks, int BFn, int H, int W) {
  pdl_launch_dependents();
  pdl_wait();
  const long long HW = static_cast<long long>(H) * W;
  const long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (idx >= BFn * HW) return;
  const int n = static_cast<int>(idx / HW);
  const long long pix = idx % HW;
  const SphereObj o = sphere_obj(circles + n * 3, H, W);
  masks[idx] = sphere_value(o, static_cast<int>(pix % W), static_cast<int>(pix / W));
}

// The scatter + unshuffle kernel with the masks generated on the fly from the circles (no mask tensor anywhere).
// Same selection rule and arithmetic as traj_unshuffle_kernel fed with sphere_mask_kernel's output (bit-identical).
constexpr int TRAJ_MAX_OBJ = 8;
__global__ void __launch_bounds__(256)
traj_circles_unshuffle_kernel(const float* __restrict__ info, const float* __restrict__ circles,
                              __nv_bfloat16* __restrict__ feat, float* __restrict__ mask_out, int BF, int n_obj, int H,
                              int W) {
  pdl_launch_dependents();
  pdl_wait();
  const int h8 = H >> 3, w8 = W >> 3;
  const long long HW = static_cast<long long>(H) * W;
  const long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  const long long total = static_cast<long long>(BF) * h8 * w8 * 8;
  if (idx >= total) return;
  const int dy = static_cast<int>(idx & 7);
  long long t = idx >> 3;
  const int cx = static_cast<int>(t % w8);
  t /= w8;
  const int cy = static_cast<int>(t % h8);
  const int n = static_cast<int>(t / h8);
  const int py = cy * 8 + dy;
  int sel[8];
  float m[8];
#pragma unroll
  for (int dx = 0; dx < 8; ++dx) {
    sel[dx] = -1;
    m[dx] = 0.f;
  }
  for (int o = 0; o < n_obj; ++o) {
    const SphereObj so = sphere_obj(circles + (static_cast<long long>(n) * n_obj + o) * 3, H, W);
#pragma unroll
    for (int dx = 0; dx < 8; ++dx) {
      const float mo = sphere_value(so, cx * 8 + dx, py);
      if (mo > 0.f) {
        sel[dx] = o;
        m[dx] = mo;
      }
    }
  }
  const long long pix0 = static_cast<long long>(py) * W + cx * 8;
  float4* mo4 = reinterpret_cast<float4*>(mask_out + static_cast<long long>(n) * HW + pix0);
  mo4[0] = make_float4(m[0], m[1], m[2], m[3]);
  mo4[1] = make_float4(m[4], m[5], m[6], m[7]);
  const float* inf = info + static_cast<long long>(n) * n_obj * 12;
  __nv_bfloat16* cell = feat + ((static_cast<long long>(n) * h8 + cy) * w8 + cx) * 832 + dy * 8;
#pragma unroll
  for (int c = 0; c < 12; ++c) {
    float r[8];
#pragma unroll
    for (int dx = 0; dx < 8; ++dx) {
      const float iv = sel[dx] >= 0 ? __ldg(inf + sel[dx] * 12 + c) : 0.f;
      r[dx] = __fmul_rn(__fmul_rn(iv, m[dx]), m[dx]);
    }
    *reinterpret_cast<uint4*>(cell + c * 64) = pack8(r);
  }
  float r[8];
#pragma unroll
  for (int dx = 0; dx < 8; ++dx) r[dx] = __fmul_rn(m[dx], m[dx]);
  *reinterpret_cast<uint4*>(cell + 12 * 64) = pack8(r);
}

// Mask modulation of an ObjectEncoder level (fmc/adapter.py:175-177): out[n, y, x, :] = x[n, y, x, :] * mask[n, ry[y], rx[x]]
// where ry / rx compose the iterated nearest-neighbour resizes down to this level (built on the host).
__global__ void __launch_bounds__(256)
mask_modulate_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ mask, const int* __restrict__ ry,
                     const int* __restrict__ rx, __nv_bfloat16* __restrict__ out, int N, int h, int w, int C, int H, int W) {
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
  const float m = __ldg(mask + (static_cast<long long>(n) * H + __ldg(ry + yy)) * W + __ldg(rx + xx));
  const long long off = ((static_cast<long long>(n) * h + yy) * w + xx) * C;
  float v[8];
  unpack8(__ldg(reinterpret_cast<const uint4*>(x + off) + vi), v);
#pragma unroll
  for (int j = 0; j < 8; ++j) v[j] *= m;
  *(reinterpret_cast<uint4*>(out + off) + vi) = pack8(v);
}

// CFG combine + DDIM (eta = 0) update on fp32 latents in the reference layout (pipeline_animation_cm_om.py:711-720,
// diffusers DDIMScheduler.step): eps = e_u + g (e_c - e_u); x0 = (x - sqrt(1-a_t) eps) / sqrt(a_t);
// x' = sqrt(a_prev) x0 + sqrt(1-a_prev) eps.   eps_c == nullptr: no guidance (eps = e_u).
__global__ void __launch_bounds__(256)
cfg_ddim_kernel(const float* __restrict__ eps_u, const float* __restrict__ eps_c, float guidance,
                const float* __restrict__ x, float* __restrict__ x_out, float* __restrict__ eps_out, float sqrt_a_t,
                float sqrt_1m_a_t, float sqrt_a_prev, float sqrt_1m_a_prev, long long n) {
  pdl_launch_dependents();
  pdl_wait();
  const long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (idx >= n) return;
  const float eu = eps_u[idx];
  const float eps = eps_c != nullptr ? eu + guidance * (eps_c[idx] - eu) : eu;
  const float x0 = (x[idx] - sqrt_1m_a_t * eps) / sqrt_a_t;
  x_out[idx] = sqrt_a_prev * x0 + sqrt_1m_a_prev * eps;
  if (eps_out != nullptr) eps_out[idx] = eps;
}

// Multidiff window average + DDIM update (fmc/pipelines/pipeline_animation.py:669-702): the long clip is denoised as
// n_windows overlapping windows of L frames (window k starts at frame k * stride); per frame the guided predictions of
// the windows covering it are averaged -- `noise_full[window] += noise_pred / count[window]` in window order, as the
// reference accumulates it -- and one DDIM (eta = 0) update is applied to the whole clip.
// eps: [n_windows, (cfg ? 2 b : b), C, L, HW] fp32 (unconditional half first); latents: [b, C, F_total, HW].
__global__ void __launch_bounds__(256)
window_combine_ddim_kernel(const float* __restrict__ eps, int n_windows, int cfg, float guidance,
                           const float* __restrict__ x, float* __restrict__ x_out, int b, int C, int F_total, long long HW,
                           int L, int stride, float sqrt_a_t, float sqrt_1m_a_t, float sqrt_a_prev, float sqrt_1m_a_prev) {
  pdl_launch_dependents();
  pdl_wait();
  const long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  const long long total = static_cast<long long>(b) * C * F_total * HW;
  if (idx >= total) return;
  const long long hw = idx % HW;
  long long t = idx / HW;
  const int f = static_cast<int>(t % F_total);
  t /= F_total;
  const int c = static_cast<int>(t % C);
  const int bi = static_cast<int>(t / C);
  // windows covering frame f: k * stride <= f < k * stride + L
  int k_lo = f - L + 1 > 0 ? (f - L + 1 + stride - 1) / stride : 0;
  int k_hi = f / stride;
  if (k_hi > n_windows - 1) k_hi = n_windows - 1;
  const float count = static_cast<float>(k_hi - k_lo + 1);
  const long long win_elems = static_cast<long long>(cfg ? 2 * b : b) * C * L * HW;
  const long long half = static_cast<long long>(b) * C * L * HW;
  float acc = 0.f;
  for (int k = k_lo; k <= k_hi; ++k) {
    const long long off = ((static_cast<long long>(bi) * C + c) * L + (f - k * stride)) * HW + hw;
    const float eu = __ldg(eps + k * win_elems + off);
    const float e = cfg ? eu + guidance * (__ldg(eps + k * win_elems + half + off) - eu) : eu;
    acc += e / count;
  }
  const float x0 = (x[idx] - sqrt_1m_a_t * acc) / sqrt_a_t;
  x_out[idx] = sqrt_a_prev * x0 + sqrt_1m_a_prev * acc;
}

static inline unsigned blocks_for(long long n) { return static_cast<unsigned>((n + 255) / 256); }

}  // namespace fmc

using namespace fmc;

extern "C" int fmc_add_bf16(const void* a, long long lda, const void* b, long long ldb, const float* rowbias,
                            int rows_per_group, long long ldrb, void* out, long long ldo, long long rows, int C,
                            int relu, void* stream) {
  FMC_REQUIRE(a && out, FMC_ERR_ARG, "fmc_add_bf16: null operand");
  FMC_REQUIRE(C % 8 == 0 && lda % 8 == 0 && ldo % 8 == 0 && (b == nullptr || ldb % 8 == 0), FMC_ERR_SHAPE,
              "fmc_add_bf16: C and strides must be multiples of 8");
  FMC_REQUIRE(rowbias == nullptr || rows_per_group > 0, FMC_ERR_ARG, "fmc_add_bf16: rows_per_group must be positive");
  if (rows == 0) return FMC_OK;
  launch_k(add_kernel, dim3(blocks_for(rows * (C / 8))), dim3(256), 0, static_cast<cudaStream_t>(stream), 
      static_cast<const __nv_bfloat16*>(a), lda, static_cast<const __nv_bfloat16*>(b), ldb, rowbias,
      rows_per_group > 0 ? rows_per_group : 1, ldrb, static_cast<__nv_bfloat16*>(out), ldo, rows, C, relu);
  return check_launch("add_kernel");
}

extern "C" int fmc_resize_nearest_bf16(const void* x, void* out, int N, int h, int w, int oh, int ow, int C,
                                       void* stream) {
  FMC_REQUIRE(x && out, FMC_ERR_ARG, "fmc_resize_nearest_bf16: null operand");
  FMC_REQUIRE(C % 8 == 0, FMC_ERR_SHAPE, "fmc_resize_nearest_bf16: C must be a multiple of 8");
  const long long total = static_cast<long long>(N) * oh * ow * (C / 8);
  if (total == 0) return FMC_OK;
  launch_k(resize_nearest_kernel, dim3(blocks_for(total)), dim3(256), 0, static_cast<cudaStream_t>(stream), 
      static_cast<const __nv_bfloat16*>(x), static_cast<__nv_bfloat16*>(out), N, h, w, oh, ow, C,
      static_cast<float>(h) / oh, static_cast<float>(w) / ow);
  return check_launch("resize_nearest_kernel");
}

extern "C" int fmc_avgpool2_bf16(const void* x, void* out, int N, int h, int w, int C, void* stream) {
  FMC_REQUIRE(x && out, FMC_ERR_ARG, "fmc_avgpool2_bf16: null operand");
  FMC_REQUIRE(C % 8 == 0, FMC_ERR_SHAPE, "fmc_avgpool2_bf16: C must be a multiple of 8");
  const long long total = static_cast<long long>(N) * (h / 2) * (w / 2) * (C / 8);
  if (total == 0) return FMC_OK;
  launch_k(avgpool2_kernel, dim3(blocks_for(total)), dim3(256), 0, static_cast<cudaStream_t>(stream), 
      static_cast<const __nv_bfloat16*>(x), static_cast<__nv_bfloat16*>(out), N, h, w, C);
  return check_launch("avgpool2_kernel");
}

extern "C" int fmc_copy2d_bf16(const void* src, long long lds, void* dst, long long ldd, long long rows, int cols,
                               void* stream) {
  FMC_REQUIRE(src && dst, FMC_ERR_ARG, "fmc_copy2d_bf16: null operand");
  FMC_REQUIRE(cols % 8 == 0 && lds % 8 == 0 && ldd % 8 == 0, FMC_ERR_SHAPE, "fmc_copy2d_bf16: cols/strides must be multiples of 8");
  if (rows == 0 || cols == 0) return FMC_OK;
  launch_k(copy2d_kernel, dim3(blocks_for(rows * (cols / 8))), dim3(256), 0, static_cast<cudaStream_t>(stream), 
      static_cast<const __nv_bfloat16*>(src), lds, static_cast<__nv_bfloat16*>(dst), ldd, rows, cols);
  return check_launch("copy2d_kernel");
}

extern "C" int fmc_ncfhw_f32_to_cl_bf16(const float* x, void* out, int B, int C, int F, long long HW, int Cpad,
                                        void* stream) {
  FMC_REQUIRE(x && out && Cpad >= C, FMC_ERR_ARG, "fmc_ncfhw_f32_to_cl_bf16: bad arguments");
  const long long total = static_cast<long long>(B) * F * HW * Cpad;
  if (total == 0) return FMC_OK;
  launch_k(ncfhw_to_cl_kernel, dim3(blocks_for(total)), dim3(256), 0, static_cast<cudaStream_t>(stream), 
      x, static_cast<__nv_bfloat16*>(out), B, C, F, HW, Cpad);
  return check_launch("ncfhw_to_cl_kernel");
}

extern "C" int fmc_cl_bf16_to_ncfhw_f32(const void* x, long long ldc, float* out, int B, int C, int F, long long HW,
                                        void* stream) {
  FMC_REQUIRE(x && out && ldc >= C, FMC_ERR_ARG, "fmc_cl_bf16_to_ncfhw_f32: bad arguments");
  const long long total = static_cast<long long>(B) * C * F * HW;
  if (total == 0) return FMC_OK;
  launch_k(cl_to_ncfhw_kernel, dim3(blocks_for(total)), dim3(256), 0, static_cast<cudaStream_t>(stream), 
      static_cast<const __nv_bfloat16*>(x), ldc, out, B, C, F, HW);
  return check_launch("cl_to_ncfhw_kernel");
}

extern "C" int fmc_cast_act_bf16(const float* x, void* out, long long n, int silu, void* stream) {
  FMC_REQUIRE(x && out, FMC_ERR_ARG, "fmc_cast_act_bf16: null operand");
  if (n == 0) return FMC_OK;
  launch_k(cast_act_kernel, dim3(blocks_for(n)), dim3(256), 0, static_cast<cudaStream_t>(stream), x, static_cast<__nv_bfloat16*>(out), n, silu);
  return check_launch("cast_act_kernel");
}

extern "C" int fmc_timestep_embedding_bf16(const float* t, void* out, int B, int dim, void* stream) {
  FMC_REQUIRE(t && out && dim % 2 == 0, FMC_ERR_ARG, "fmc_timestep_embedding_bf16: bad arguments");
  if (B == 0) return FMC_OK;
  launch_k(timestep_embedding_kernel, dim3((B * dim / 2 + 127) / 128), dim3(128), 0, static_cast<cudaStream_t>(stream), 
      t, static_cast<__nv_bfloat16*>(out), B, dim);
  return check_launch("timestep_embedding_kernel");
}

extern "C" int fmc_plucker_f32(const float* K, const float* c2w, float* out, int BF, int H, int W, void* stream) {
  FMC_REQUIRE(K && c2w && out, FMC_ERR_ARG, "fmc_plucker_f32: null operand");
  const long long total = static_cast<long long>(BF) * H * W;
  if (total == 0) return FMC_OK;
  launch_k(plucker_plain_kernel, dim3(blocks_for(total)), dim3(256), 0, static_cast<cudaStream_t>(stream), K, c2w, out, BF, H, W);
  return check_launch("plucker_plain_kernel");
}

extern "C" int fmc_plucker_unshuffle_bf16(const float* K, const float* c2w, void* out, int BF, int H, int W,
                                          void* stream) {
  FMC_REQUIRE(K && c2w && out, FMC_ERR_ARG, "fmc_plucker_unshuffle_bf16: null operand");
  FMC_REQUIRE(H % 8 == 0 && W % 8 == 0, FMC_ERR_SHAPE, "fmc_plucker_unshuffle_bf16: H=%d W=%d must be multiples of 8", H, W);
  const long long total = static_cast<long long>(BF) * (H / 8) * (W / 8) * 8;
  if (total == 0) return FMC_OK;
  launch_k(plucker_unshuffle_kernel, dim3(blocks_for(total)), dim3(256), 0, static_cast<cudaStream_t>(stream), 
      K, c2w, static_cast<__nv_bfloat16*>(out), BF, H, W);
  return check_launch("plucker_unshuffle_kernel");
}

extern "C" int fmc_traj_scatter_f32(const float* info, const float* masks, float* feat, float* mask_out, int BF,
                                    int n_obj, int H, int W, void* stream) {
  FMC_REQUIRE(info && masks && feat && mask_out, FMC_ERR_ARG, "fmc_traj_scatter_f32: null operand");
  const long long total = static_cast<long long>(BF) * H * W;
  if (total == 0) return FMC_OK;
  launch_k(traj_plain_kernel, dim3(blocks_for(total)), dim3(256), 0, static_cast<cudaStream_t>(stream), info, masks, feat, mask_out, BF, n_obj, H, W);
  return check_launch("traj_plain_kernel");
}

extern "C" int fmc_traj_scatter_unshuffle_bf16(const float* info, const float* masks, void* feat, float* mask_out,
                                               int BF, int n_obj, int H, int W, void* stream) {
  FMC_REQUIRE(info && masks && feat && mask_out, FMC_ERR_ARG, "fmc_traj_scatter_unshuffle_bf16: null operand");
  FMC_REQUIRE(H % 8 == 0 && W % 8 == 0, FMC_ERR_SHAPE, "fmc_traj_scatter_unshuffle_bf16: H=%d W=%d must be multiples of 8", H, W);
  const long long total = static_cast<long long>(BF) * (H / 8) * (W / 8) * 8;
  if (total == 0) return FMC_OK;
  launch_k(traj_unshuffle_kernel, dim3(blocks_for(total)), dim3(256), 0, static_cast<cudaStream_t>(stream), 
      info, masks, static_cast<__nv_bfloat16*>(feat), mask_out, BF, n_obj, H, W);
  return check_launch("traj_unshuffle_kernel");
}

extern "C" int fmc_mask_modulate_bf16(const void* x, const float* mask, const int* row_index, const int* col_index,
                                      void* out, int N, int h, int w, int C, int H, int W, void* stream) {
  FMC_REQUIRE(x && mask && row_index && col_index && out, FMC_ERR_ARG, "fmc_mask_modulate_bf16: null operand");
  FMC_REQUIRE(C % 8 == 0, FMC_ERR_SHAPE, "fmc_mask_modulate_bf16: C must be a multiple of 8");
  const long long total = static_cast<long long>(N) * h * w * (C / 8);
  if (total == 0) return FMC_OK;
  launch_k(mask_modulate_kernel, dim3(blocks_for(total)), dim3(256), 0, static_cast<cudaStream_t>(stream), 
      static_cast<const __nv_bfloat16*>(x), mask, row_index, col_index, static_cast<__nv_bfloat16*>(out), N, h, w, C, H, W);
  return check_launch("mask_modulate_kernel");
}

extern "C" int fmc_cfg_ddim_step_f32(const float* eps_uncond, const float* eps_cond, float guidance_scale,
                                     const float* latents, float* latents_out, float* eps_out, float alpha_t,
                                     float alpha_prev, long long n, void* stream) {
  FMC_REQUIRE(eps_uncond && latents && latents_out, FMC_ERR_ARG, "fmc_cfg_ddim_step_f32: null operand");
  FMC_REQUIRE(alpha_t > 0.f && alpha_t <= 1.f && alpha_prev > 0.f && alpha_prev <= 1.f, FMC_ERR_ARG,
              "fmc_cfg_ddim_step_f32: alphas must be in (0, 1]");
  if (n == 0) return FMC_OK;
  launch_k(cfg_ddim_kernel, dim3(blocks_for(n)), dim3(256), 0, static_cast<cudaStream_t>(stream), 
      eps_uncond, eps_cond, guidance_scale, latents, latents_out, eps_out, sqrtf(alpha_t), sqrtf(1.f - alpha_t),
      sqrtf(alpha_prev), sqrtf(1.f - alpha_prev), n);
  return check_launch("cfg_ddim_kernel");
}

extern "C" int fmc_window_combine_ddim_f32(const float* eps_windows, int n_windows, int cfg, float guidance_scale,
                                           const float* latents, float* latents_out, int b, int C, int F_total,
                                           long long HW, int L, int stride, float alpha_t, float alpha_prev,
                                           void* stream) {
  FMC_REQUIRE(eps_windows && latents && latents_out, FMC_ERR_ARG, "fmc_window_combine_ddim_f32: null operand");
  FMC_REQUIRE(n_windows >= 1 && L >= 1 && stride >= 1 && stride <= L && F_total == (n_windows - 1) * stride + L, FMC_ERR_SHAPE,
              "fmc_window_combine_ddim_f32: %d windows of %d frames with stride %d do not tile %d frames", n_windows, L,
              stride, F_total);
  FMC_REQUIRE(alpha_t > 0.f && alpha_t <= 1.f && alpha_prev > 0.f && alpha_prev <= 1.f, FMC_ERR_ARG,
              "fmc_window_combine_ddim_f32: alphas must be in (0, 1]");
  const long long total = static_cast<long long>(b) * C * F_total * HW;
  if (total == 0) return FMC_OK;
  launch_k(window_combine_ddim_kernel, dim3(blocks_for(total)), dim3(256), 0, static_cast<cudaStream_t>(stream), eps_windows,
           n_windows, cfg, guidance_scale, latents, latents_out, b, C, F_total, HW, L, stride, sqrtf(alpha_t),
           sqrtf(1.f - alpha_t), sqrtf(alpha_prev), sqrtf(1.f - alpha_prev));
  return check_launch("window_combine_ddim_kernel");
}

extern "C" int fmc_sphere_mask_f32(const float* circles, float* masks, int BF, int n_obj, int H, int W, void* stream) {
  FMC_REQUIRE(circles && masks, FMC_ERR_ARG, "fmc_sphere_mask_f32: null operand");
  const long long total = static_cast<long long>(BF) * n_obj * H * W;
  if (total == 0) return FMC_OK;
  FMC_REQUIRE(static_cast<long long>(BF) * n_obj < (1LL << 31), FMC_ERR_SHAPE, "fmc_sphere_mask_f32: too many objects");
  launch_k(sphere_mask_kernel, dim3(blocks_for(total)), dim3(256), 0, static_cast<cudaStream_t>(stream), circles, masks,
           BF * n_obj, H, W);
  return check_launch("sphere_mask_kernel");
}

extern "C" int fmc_traj_scatter_circles_unshuffle_bf16(const float* info, const float* circles, void* feat,
                                                       float* mask_out, int BF, int n_obj, int H, int W, void* stream) {
  FMC_REQUIRE(info && circles && feat && mask_out, FMC_ERR_ARG, "fmc_traj_scatter_circles_unshuffle_bf16: null operand");
  FMC_REQUIRE(H % 8 == 0 && W % 8 == 0, FMC_ERR_SHAPE, "fmc_traj_scatter_circles_unshuffle_bf16: H=%d W=%d must be multiples of 8", H, W);
  const long long total = static_cast<long long>(BF) * (H / 8) * (W / 8) * 8;
  if (total == 0) return FMC_OK;
  launch_k(traj_circles_unshuffle_kernel, dim3(blocks_for(total)), dim3(256), 0, static_cast<cudaStream_t>(stream), info,
           circles, static_cast<__nv_bfloat16*>(feat), mask_out, BF, n_obj, H, W);
  return check_launch("traj_circles_unshuffle_kernel");
}
