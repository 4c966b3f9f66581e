// Host-side helpers shared by the C-ABI entry points: thread-local error string,
// CUDA error mapping, and TMA tensor-map construction through the driver entry
// point (no link-time dependency on libcuda).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>

#include "burst_attn_b200.h"

namespace ba {

void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what);

#define BA_CHECK_CUDA(expr)                                  \
  do {                                                       \
    cudaError_t _e = (expr);                                 \
    if (_e != cudaSuccess) return ::ba::cuda_fail(_e, #expr); \
  } while (0)

#define BA_REQUIRE(cond, ...)          \
  do {                                 \
    if (!(cond)) {                     \
      ::ba::set_error(__VA_ARGS__);    \
      return BA_ERR_INVALID;           \
    }                                  \
  } while (0)

// Build a 4-D tiled tensor map over a [b,s,h,d] view (d contiguous):
// dims innermost-first (D, H, S, B); box (box_d, 1, box_s, 1).
// esize: element size in bytes; swizzle128: CU_TENSOR_MAP_SWIZZLE_128B when
// box_d*esize == 128, else no swizzle.
int make_tensor_map(CUtensorMap* out, const ba_tensor4& t, int B, int S, int H, int D, CUtensorMapDataType dt,
                    int esize, int box_d, int box_s, bool swizzle128);

inline CUtensorMapDataType lowp_dtype(int dtype) {
  return dtype == BA_DTYPE_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
}

}  // namespace ba
