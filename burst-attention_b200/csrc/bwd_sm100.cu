#include "host_common.h"
extern "C" int ba_bwd_chunk(ba_tensor4 d_o, ba_tensor4 q, ba_tensor4 k, ba_tensor4 v, ba_rowstat delta, ba_rowstat lse,
                 ba_tensor4 dq_acc, ba_tensor4 dk_acc, ba_tensor4 dv_acc, int B, int Sq, int Sk, int H, int D,
                 float scale, int mask_mode, int causal_offset, int flags, int dtype, void* stream) {
  ba::set_error("ba_bwd_chunk: not built yet");
  return BA_ERR_UNSUPPORTED;
}
