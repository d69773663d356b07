// libadvgrpo_b200: ABI bookkeeping (version, per-thread error string, device check) and
// the host-side TMA tensor-map encoder shared by the tensor-core kernels.
#include <stdarg.h>

#include <stdlib.h>

#include "common.cuh"

namespace advgrpo {

thread_local char g_last_error[512] = {0};

int set_error(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_last_error, sizeof(g_last_error), fmt, ap);
  va_end(ap);
  return code;
}

int sm_count() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached[dev] = n;
  }
  return cached[dev];
}

static int g_pdl = -1;   // -1: read ADVGRPO_PDL on first use
bool pdl_enabled() {
  if (g_pdl < 0) {
    const char* e = getenv("ADVGRPO_PDL");
    g_pdl = (e && e[0] == '1') ? 1 : 0;
  }
  return g_pdl == 1;
}
void set_pdl(int v) { g_pdl = v ? 1 : 0; }

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn) return fn;
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
      q != cudaDriverEntryPointSuccess)
    return nullptr;
  fn = reinterpret_cast<EncodeTiledFn>(p);
  return fn;
}

int make_tmap_bf16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims,
                   const uint64_t* strides_bytes, const uint32_t* box, bool swizzle128) {
  return make_tmap(out, base, rank, dims, strides_bytes, box, swizzle128, false);
}

int make_tmap(CUtensorMap* out, const void* base, int rank, const uint64_t* dims,
              const uint64_t* strides_bytes, const uint32_t* box, bool swizzle128, bool f32) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return set_error(ADVGRPO_ERR_CUDA, "cuTensorMapEncodeTiled entry point unavailable");
  cuuint64_t gdim[5];
  cuuint64_t gstr[5];
  cuuint32_t bx[5];
  cuuint32_t estr[5];
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bx[i] = box[i];
    estr[i] = 1;
    if (i > 0) gstr[i - 1] = strides_bytes[i];
  }
  CUresult r = fn(out, f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base),
                  gdim, gstr, bx, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return set_error(ADVGRPO_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  return ADVGRPO_OK;
}

}  // namespace advgrpo

extern "C" {

int advgrpo_abi_version(void) { return ADVGRPO_ABI_VERSION; }

const char* advgrpo_last_error(void) { return advgrpo::g_last_error; }

// Test/bench hook: programmatic dependent launch for the MMDiT kernel chain (see common.cuh).
void advgrpo_debug_set_pdl(int v) { advgrpo::set_pdl(v); }

int advgrpo_device_check(int dev) {
  cudaDeviceProp p;
  cudaError_t e = cudaGetDeviceProperties(&p, dev);
  if (e != cudaSuccess)
    return advgrpo::set_error(ADVGRPO_ERR_CUDA, "cudaGetDeviceProperties(%d): %s", dev,
                              cudaGetErrorString(e));
  if (p.major != 10)
    return advgrpo::set_error(ADVGRPO_ERR_UNSUPPORTED,
                              "device %d is sm_%d%d; this library is built for sm_100a only", dev,
                              p.major, p.minor);
  return ADVGRPO_OK;
}

}  // extern "C"
