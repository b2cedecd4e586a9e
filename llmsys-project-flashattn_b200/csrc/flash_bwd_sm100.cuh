// Flash-attention backward for sm_100a (tcgen05 + TMEM).  Placeholder until the tensor-core
// backward lands: reports FA_ERR_UNSUPPORTED so the caller uses the CUDA-core backward.
#pragma once
#include "ptx.cuh"

namespace fa {
namespace sm100 {

inline int bwd_tc(const fa_attn_desc*, const void*, const void*, const void*, const void*, const void*, const float*,
                  const float*, void*, void*, void*, cudaStream_t) {
  return FA_ERR_UNSUPPORTED;
}

}  // namespace sm100
}  // namespace fa
