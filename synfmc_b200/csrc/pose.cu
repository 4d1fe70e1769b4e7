// Relative-pose algebra of the data pipeline on the device: the three functions of fmc/data/utils.py:148-200
// that turn absolute camera / object poses into the relative 3x4 blocks feeding the Pluecker embedding (a9) and the
// ObjectEncoder pose features (a10), so poses that already live on the GPU never return to numpy.  One thread per pose,
// fp64 like the reference (numpy on float64 arrays); a pose is a row-major 3x4 block [R | T] at the head of a `stride`-double
// record (12 for 3x4 storage, 16 for 4x4).
#include "common.cuh"
#include "ptx.cuh"

namespace fmc {

struct Pose {
  double r[3][3];
  double t[3];
};
__device__ __forceinline__ Pose load_pose(const double* p) {
  Pose q;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
#pragma unroll
    for (int j = 0; j < 3; ++j) q.r[i][j] = p[i * 4 + j];
    q.t[i] = p[i * 4 + 3];
  }
  return q;
}
// out = [R^T R_ref | R^T (T_ref - T_sub) / scale]   (R from `a`; T_sub is a's own translation unless another is given)
__device__ __forceinline__ void store_relative(const Pose& a, const double (&t_sub)[3], const Pose& ref, double scale, double* out) {
#pragma unroll
  for (int i = 0; i < 3; ++i) {
#pragma unroll
    for (int j = 0; j < 3; ++j) out[i * 4 + j] = a.r[0][i] * ref.r[0][j] + a.r[1][i] * ref.r[1][j] + a.r[2][i] * ref.r[2][j];
    const double neg = -(a.r[0][i] * t_sub[0] + a.r[1][i] * t_sub[1] + a.r[2][i] * t_sub[2]);
    const double pos = a.r[0][i] * ref.t[0] + a.r[1][i] * ref.t[1] + a.r[2][i] * ref.t[2];
    out[i * 4 + 3] = (neg + pos) / scale;
  }
}

// create_relative_matrix_of_cam_list (fmc/data/utils.py:148-163): poses of a clip relative to its first frame; frame 0 is
// exactly eye(3, 4)
__global__ void pose_relative_to_first_kernel(const double* __restrict__ poses, long long stride, double* __restrict__ out,
                                              int clips, int frames, double scale) {
  pdl_launch_dependents();
  pdl_wait();
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= clips * frames) return;
  const int clip = idx / frames, f = idx % frames;
  double* o = out + static_cast<long long>(idx) * 12;
  if (f == 0) {
#pragma unroll
    for (int i = 0; i < 12; ++i) o[i] = (i == 0 || i == 5 || i == 10) ? 1.0 : 0.0;
    return;
  }
  const Pose first = load_pose(poses + static_cast<long long>(clip) * frames * stride);
  const Pose me = load_pose(poses + static_cast<long long>(idx) * stride);
  store_relative(me, me.t, first, scale, o);
}

// create_absolute_matrix_from_ref_cam_list (fmc/data/utils.py:167-183): first @ inv([rel with T * scale; 0 0 0 1]), rows
// 0..2; frame 0 = the first pose itself.  The inverse is the general affine one (adjugate of the 3x3 block), as
// np.linalg.inv does not assume a rotation.
__global__ void pose_absolute_from_relative_kernel(const double* __restrict__ first, const double* __restrict__ rel,
                                                   double* __restrict__ out, int clips, int frames, double scale) {
  pdl_launch_dependents();
  pdl_wait();
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= clips * frames) return;
  const int clip = idx / frames, f = idx % frames;
  const double* F = first + static_cast<long long>(clip) * 16;
  double* o = out + static_cast<long long>(idx) * 12;
  if (f == 0) {
#pragma unroll
    for (int i = 0; i < 12; ++i) o[i] = F[i];
    return;
  }
  const Pose q = load_pose(rel + static_cast<long long>(idx) * 12);
  const double t[3] = {q.t[0] * scale, q.t[1] * scale, q.t[2] * scale};
  const double (&a)[3][3] = q.r;
  double inv[3][3];
  inv[0][0] = a[1][1] * a[2][2] - a[1][2] * a[2][1];
  inv[0][1] = a[0][2] * a[2][1] - a[0][1] * a[2][2];
  inv[0][2] = a[0][1] * a[1][2] - a[0][2] * a[1][1];
  inv[1][0] = a[1][2] * a[2][0] - a[1][0] * a[2][2];
  inv[1][1] = a[0][0] * a[2][2] - a[0][2] * a[2][0];
  inv[1][2] = a[0][2] * a[1][0] - a[0][0] * a[1][2];
  inv[2][0] = a[1][0] * a[2][1] - a[1][1] * a[2][0];
  inv[2][1] = a[0][1] * a[2][0] - a[0][0] * a[2][1];
  inv[2][2] = a[0][0] * a[1][1] - a[0][1] * a[1][0];
  const double det = a[0][0] * inv[0][0] + a[0][1] * inv[1][0] + a[0][2] * inv[2][0];
  double it[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) {
#pragma unroll
    for (int j = 0; j < 3; ++j) inv[i][j] /= det;
  }
#pragma unroll
  for (int i = 0; i < 3; ++i) it[i] = -(inv[i][0] * t[0] + inv[i][1] * t[1] + inv[i][2] * t[2]);
#pragma unroll
  for (int i = 0; i < 3; ++i) {
#pragma unroll
    for (int j = 0; j < 3; ++j) o[i * 4 + j] = F[i * 4 + 0] * inv[0][j] + F[i * 4 + 1] * inv[1][j] + F[i * 4 + 2] * inv[2][j];
    o[i * 4 + 3] = F[i * 4 + 0] * it[0] + F[i * 4 + 1] * it[1] + F[i * 4 + 2] * it[2] + F[i * 4 + 3];
  }
}

// create_relative_matrix_of_two_torch_matrix (fmc/data/utils.py:185-200): n object poses relative to one camera pose per
// set.  Reference behaviour kept: its stacked np.dot(...)[..., 0, 0] takes the subtracted translation from OBJECT 0 of the
// set for every object (identical to the intended formula when a set has one object).
__global__ void pose_objects_relative_kernel(const double* __restrict__ cam, long long cam_stride, const double* __restrict__ obj,
                                             long long obj_stride, double* __restrict__ out, int sets, int n, double scale) {
  pdl_launch_dependents();
  pdl_wait();
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= sets * n) return;
  const int set = idx / n;
  const Pose c = load_pose(cam + static_cast<long long>(set) * cam_stride);
  const Pose me = load_pose(obj + static_cast<long long>(idx) * obj_stride);
  const Pose o0 = load_pose(obj + static_cast<long long>(set) * n * obj_stride);
  store_relative(me, o0.t, c, scale, out + static_cast<long long>(idx) * 12);
}

}  // namespace fmc

using namespace fmc;

extern "C" int fmc_pose_relative_to_first_f64(const double* poses, long long pose_stride, double* out, int clips, int frames,
                                              double scale_T, void* stream) {
  FMC_REQUIRE(poses && out && pose_stride >= 12 && scale_T != 0.0, FMC_ERR_ARG, "fmc_pose_relative_to_first_f64: bad arguments");
  if (clips * frames == 0) return FMC_OK;
  launch_k(pose_relative_to_first_kernel, dim3((clips * frames + 127) / 128), dim3(128), 0, static_cast<cudaStream_t>(stream),
           poses, pose_stride, out, clips, frames, scale_T);
  return check_launch("pose_relative_to_first_kernel");
}

extern "C" int fmc_pose_absolute_from_relative_f64(const double* first, const double* rel, double* out, int clips, int frames,
                                                   double scale_T, void* stream) {
  FMC_REQUIRE(first && rel && out, FMC_ERR_ARG, "fmc_pose_absolute_from_relative_f64: null operand");
  if (clips * frames == 0) return FMC_OK;
  launch_k(pose_absolute_from_relative_kernel, dim3((clips * frames + 127) / 128), dim3(128), 0,
           static_cast<cudaStream_t>(stream), first, rel, out, clips, frames, scale_T);
  return check_launch("pose_absolute_from_relative_kernel");
}

extern "C" int fmc_pose_objects_relative_f64(const double* cam, long long cam_stride, const double* obj, long long obj_stride,
                                             double* out, int sets, int n, double scale_T, void* stream) {
  FMC_REQUIRE(cam && obj && out && cam_stride >= 12 && obj_stride >= 12 && scale_T != 0.0, FMC_ERR_ARG,
              "fmc_pose_objects_relative_f64: bad arguments");
  if (sets * n == 0) return FMC_OK;
  launch_k(pose_objects_relative_kernel, dim3((sets * n + 127) / 128), dim3(128), 0, static_cast<cudaStream_t>(stream), cam,
           cam_stride, obj, obj_stride, out, sets, n, scale_T);
  return check_launch("pose_objects_relative_kernel");
}
