#include "host_common.h"

#include <mutex>
#include <string.h>

namespace ba {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int cuda_fail(cudaError_t e, const char* what) {
  set_error("CUDA error %d (%s) at %s", static_cast<int>(e), cudaGetErrorString(e), what);
  return BA_ERR_CUDA;
}

typedef CUresult (*encode_tiled_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static encode_tiled_fn get_encode() {
  static encode_tiled_fn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess) {
      fn = reinterpret_cast<encode_tiled_fn>(p);
    }
  });
  return fn;
}

int make_tensor_map(CUtensorMap* out, const ba_tensor4& t, int B, int S, int H, int D, CUtensorMapDataType dt,
                    int esize, int box_d, int box_s, bool swizzle128) {
  encode_tiled_fn enc = get_encode();
  if (!enc) {
    set_error("cuTensorMapEncodeTiled driver entry point not available");
    return BA_ERR_CUDA;
  }
  if ((reinterpret_cast<uintptr_t>(t.ptr) & 15) != 0) {
    set_error("tensor base pointer must be 16-byte aligned for TMA");
    return BA_ERR_INVALID;
  }
  // A dimension of extent 1 never contributes to an address; give it a legal stride.
  int64_t sh = (H == 1) ? D : t.stride_h;
  int64_t ss = (S == 1) ? (int64_t)D * H : t.stride_s;
  int64_t sb = (B == 1) ? ss * S : t.stride_b;
  if (B == 1 && sb < (int64_t)D) sb = (int64_t)D * H * S;
  cuuint64_t dims[4] = {(cuuint64_t)D, (cuuint64_t)H, (cuuint64_t)S, (cuuint64_t)B};
  cuuint64_t strides[3] = {(cuuint64_t)sh * esize, (cuuint64_t)ss * esize, (cuuint64_t)sb * esize};
  for (int i = 0; i < 3; ++i) {
    if (strides[i] % 16 != 0 || strides[i] == 0) {
      set_error("tensor stride %d (= %llu bytes) must be a positive multiple of 16 bytes for TMA", i,
                (unsigned long long)strides[i]);
      return BA_ERR_INVALID;
    }
  }
  cuuint32_t box[4] = {(cuuint32_t)box_d, 1u, (cuuint32_t)box_s, 1u};
  cuuint32_t estr[4] = {1u, 1u, 1u, 1u};
  CUresult r = enc(out, dt, 4, t.ptr, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with CUresult %d (dims %d,%d,%d,%d box %d,%d)", (int)r, D, H, S, B,
              box_d, box_s);
    return BA_ERR_CUDA;
  }
  return BA_OK;
}

}  // namespace ba

#ifdef BA_SELFTEST_LIB
extern "C" const char* ba_selftest_last_error(void) { return ba::g_err; }
#else
extern "C" const char* ba_last_error(void) { return ba::g_err; }
extern "C" int ba_version(void) { return 200; }
extern "C" int ba_device_check(void) {
  int dev = 0;
  BA_CHECK_CUDA(cudaGetDevice(&dev));
  int major = 0, minor = 0;
  BA_CHECK_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  BA_CHECK_CUDA(cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev));
  if (major != 10) {
    ba::set_error("burst_attn_b200 needs an sm_100 (B200) device, found sm_%d%d", major, minor);
    return BA_ERR_UNSUPPORTED;
  }
  return BA_OK;
}
#endif
