#include "common.cuh"

#include <cstdarg>
#include <cstdlib>
#include <cstring>
#include <mutex>

namespace fmc {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

bool pdl_enabled() {
  static const bool on = [] {
    const char* e = getenv("FMC_PDL");
    return e != nullptr && e[0] == '1';
  }();
  return on;
}

int current_device_ordinal() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0;
  return dev;
}

int device_sm_count() {
  static int cached[64] = {0};
  const int dev = current_device_ordinal() & 63;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) return 148;
    cached[dev] = n;
  }
  return cached[dev];
}

bool first_use_on_this_device(unsigned long long* seen_mask) {
  const unsigned long long bit = 1ull << (current_device_ordinal() & 63);
  if (*seen_mask & bit) return false;
  *seen_mask |= bit;
  return true;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess) {
      fn = reinterpret_cast<EncodeTiledFn>(p);
    }
  });
  return fn;
}

int make_tmap_bf16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                   const uint32_t* box, bool swizzle128) {
  return make_tmap_bf16_sw(out, base, rank, dims, strides_bytes, box, swizzle128 ? 128 : 0);
}

int make_tmap_bf16_sw(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                      const uint32_t* box, int swizzle_bytes) {
  const uint32_t ones[5] = {1, 1, 1, 1, 1};
  return make_tmap_bf16_ex(out, base, rank, dims, strides_bytes, box, ones, swizzle_bytes);
}

static int make_tmap_typed(CUtensorMap* out, CUtensorMapDataType dtype, const void* base, int rank, const uint64_t* dims,
                           const uint64_t* strides_bytes, const uint32_t* box, const uint32_t* estr_in, int swizzle_bytes);

int make_tmap_bf16_ex(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                      const uint32_t* box, const uint32_t* estr_in, int swizzle_bytes) {
  return make_tmap_typed(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, base, rank, dims, strides_bytes, box, estr_in, swizzle_bytes);
}

int make_tmap_f32(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                  const uint32_t* box, int swizzle_bytes) {
  static const bool tf = [] {
    const char* e = getenv("FMC_TF32_TMA");
    return e != nullptr && strcmp(e, "tfloat32") == 0;
  }();
  const uint32_t ones[5] = {1, 1, 1, 1, 1};
  return make_tmap_typed(out, tf ? CU_TENSOR_MAP_DATA_TYPE_TFLOAT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, base, rank, dims,
                         strides_bytes, box, ones, swizzle_bytes);
}

static int make_tmap_typed(CUtensorMap* out, CUtensorMapDataType dtype, const void* base, int rank, const uint64_t* dims,
                           const uint64_t* strides_bytes, const uint32_t* box, const uint32_t* estr_in, int swizzle_bytes) {
  EncodeTiledFn fn = get_encode_fn();
  FMC_REQUIRE(fn != nullptr, FMC_ERR_DRIVER, "cuTensorMapEncodeTiled entry point not available");
  FMC_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0, FMC_ERR_SHAPE, "TMA base %p not 16-byte aligned", base);
  cuuint64_t gdim[5];
  cuuint64_t gstr[5];
  cuuint32_t bdim[5];
  cuuint32_t estr[5];
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bdim[i] = box[i];
    estr[i] = estr_in[i];
    if (i > 0) {
      gstr[i - 1] = strides_bytes[i - 1];
      FMC_REQUIRE((gstr[i - 1] & 15) == 0, FMC_ERR_SHAPE, "TMA stride %llu not a multiple of 16 bytes",
                  (unsigned long long)gstr[i - 1]);
    }
  }
  CUresult r = fn(out, dtype, static_cast<cuuint32_t>(rank), const_cast<void*>(base), gdim,
                  gstr, bdim, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  swizzle_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                  : swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                  : swizzle_bytes == 32 ? CU_TENSOR_MAP_SWIZZLE_32B : CU_TENSOR_MAP_SWIZZLE_NONE,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  FMC_REQUIRE(r == CUDA_SUCCESS, FMC_ERR_DRIVER, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  return FMC_OK;
}

}  // namespace fmc

extern "C" const char* fmc_last_error_string(void) { return fmc::g_err; }

extern "C" int fmc_abi_version(void) { return FMC_B200_ABI_VERSION; }
