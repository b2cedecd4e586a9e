// Host-side TMA descriptor helper shared by the libraries that launch tcgen05 kernels.  cuTensorMapEncodeTiled is
// fetched through the runtime (cudaGetDriverEntryPoint), so nothing links against libcuda.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

namespace fa {

using EncodeTiledFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// 2-D map over a row-major matrix: `cols` contiguous elements per row, `rows` rows `ld` elements apart; box =
// [box_rows][box_cols], 128-byte swizzle, out-of-bounds elements read as zero / are not written.
static int make_tmap_2d(CUtensorMap* tm, const void* base, CUtensorMapDataType dt, int esize, long long cols,
                        long long rows, long long ld, int box_cols, int box_rows) {
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) return set_error(FA_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * esize};
  cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(tm, dt, 2, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return set_error(FA_ERR_CUDA, "cuTensorMapEncodeTiled (2-D) failed (CUresult %d; base %p cols %lld rows %lld ld %lld)",
                     (int)r, base, cols, rows, ld);
  return FA_OK;
}

}  // namespace fa
