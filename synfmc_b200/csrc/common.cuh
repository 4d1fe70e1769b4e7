// Host-side plumbing shared by all launchers of libfmc_b200: error reporting for the C ABI and
// TMA tensor-map construction (driver entry point resolved at run time so the library loads,
// and its symbols can be inspected, on a machine without a GPU driver).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <utility>

#include "../../include/fmc_b200.h"

namespace fmc {

// Error codes returned through the C ABI (0 = success).
enum : int {
  FMC_OK = 0,
  FMC_ERR_SHAPE = -1,    // unsupported shape / alignment for the kernel
  FMC_ERR_CUDA = -2,     // CUDA runtime error at launch
  FMC_ERR_DRIVER = -3,   // driver entry point / tensor-map encode failure
  FMC_ERR_ARG = -4,      // null pointer or inconsistent arguments
};

void set_error(const char* fmt, ...);

#define FMC_REQUIRE(cond, code, ...) \
  do {                               \
    if (!(cond)) {                   \
      ::fmc::set_error(__VA_ARGS__); \
      return (code);                 \
    }                                \
  } while (0)

#define FMC_CUDA_OK(expr)                                                               \
  do {                                                                                  \
    cudaError_t _e = (expr);                                                            \
    if (_e != cudaSuccess) {                                                            \
      ::fmc::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, \
                       __LINE__);                                                       \
      return ::fmc::FMC_ERR_CUDA;                                                       \
    }                                                                                   \
  } while (0)

inline int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("launch of %s failed: %s", what, cudaGetErrorString(e));
    return FMC_ERR_CUDA;
  }
  return FMC_OK;
}

int device_sm_count();  // of the CURRENT device (cached per device ordinal)

// true exactly once per (call site, current device): guards cudaFuncSetAttribute calls, which apply per device
// (a process may drive several GPUs; one process per GPU is the normal deployment)
bool first_use_on_this_device(unsigned long long* seen_mask);
int current_device_ordinal();

// Programmatic dependent launch: every kernel of this library starts with pdl_wait() (griddepcontrol.wait: all memory
// of the preceding kernel is visible once it returns) and is launched with programmatic stream serialisation allowed,
// so its CTAs may be scheduled -- barrier init, TMEM allocation, descriptor prefetch -- while the tail of the
// preceding kernel drains (every kernel triggers its dependents at its first instruction and waits after its own
// prologue).  Measured on the graph-replayed step (round 1, two alternating runs each): 31.58 / 31.27 ms with,
// 31.39 / 31.42 ms without -- no difference, so it is OFF unless FMC_PDL=1 (both instructions are no-ops for a kernel
// launched without the attribute).
bool pdl_enabled();

template <typename... KArgs, typename... Args>
inline cudaError_t launch_k(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                            Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}

// bf16 tensor map, up to 4 dims; dims[0]/box[0] innermost (elements), strides[i] in BYTES for dim i+1.
// swizzle128: CU_TENSOR_MAP_SWIZZLE_128B (inner box must then be <= 64 bf16).
int make_tmap_bf16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                   const uint32_t* box, bool swizzle128);

// same with an explicit swizzle span in bytes (0, 32, 64 or 128)
int make_tmap_bf16_sw(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                      const uint32_t* box, int swizzle_bytes);

// same with per-dimension element (traversal) strides: box[i] elements are traversed, every estr[i]-th one is copied
int make_tmap_bf16_ex(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                      const uint32_t* box, const uint32_t* estr, int swizzle_bytes);

// fp32 tensor map (reference-precision mode, csrc/precise.cu); swizzle128: inner box must be <= 32 floats.
// FMC_TF32_TMA=tfloat32 encodes CU_TENSOR_MAP_DATA_TYPE_TFLOAT32 instead of FLOAT32 (diagnostic switch).
int make_tmap_f32(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                  const uint32_t* box, int swizzle_bytes);

static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

}  // namespace fmc
