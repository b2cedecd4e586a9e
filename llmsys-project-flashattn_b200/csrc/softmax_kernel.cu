// softmax_kernel.so -- masked attention softmax forward / backward for sm_100a.
//
// Replaces the reference's LightSeq-derived src/softmax_kernel.cu (ker_attn_softmax_lt32 :36,
// ker_attn_softmax :125, ker_attn_softmax_bw :309) behind the same C ABI
// (launch_attn_softmax :233, launch_attn_softmax_bw :345).  These ops are HBM-bound
// (fw 8 B/elem, bw 12 B/elem): every row is read exactly once with 128-bit coalesced loads,
// kept in registers for the max / sum / normalise passes, and written once with 128-bit
// stores.  A row is owned by one warp (row stats by __shfl_xor) up to 2048 columns and by one
// 256-thread CTA above that; the reference's to_len<=1024 / <=2048 caps are lifted.
#include <cfloat>
#include <cstdint>

#include "host_common.cuh"

namespace fa {

constexpr float kSoftmaxEps = 1e-8f;          // reference src/softmax_kernel.cu:12
constexpr float kMaskedScore = -100000000.f;  // REDUCE_FLOAT_INF_NEG, src/includes/block_reduce.h:13

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Reduction over the TPR threads that own one row (TPR = 32: a warp shuffle; otherwise TPR/32
// warps of the 256-thread CTA combine through shared memory; every thread of the CTA must call it).
template <bool IS_MAX, int TPR>
__device__ __forceinline__ float row_reduce(float v, float* red) {
  v = IS_MAX ? warp_max(v) : warp_sum(v);
  if constexpr (TPR == 32) {
    return v;
  } else {
    constexpr int WPG = TPR / 32;  // warps per row group
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    __syncthreads();  // protect `red` from the previous use
    if (lane == 0) red[w] = v;
    __syncthreads();
    const int g0 = (w / WPG) * WPG;
    float r = red[g0];
#pragma unroll
    for (int i = 1; i < WPG; ++i) r = IS_MAX ? fmaxf(r, red[g0 + i]) : r + red[g0 + i];
    return r;
  }
}

template <int VEC>
struct Vec;
template <>
struct Vec<4> {
  using T = float4;
};
template <>
struct Vec<1> {
  using T = float;
};

template <int VEC>
__device__ __forceinline__ void load_vec(const float* p, float (&v)[VEC]) {
  if constexpr (VEC == 4) {
    float4 t = __ldg(reinterpret_cast<const float4*>(p));
    v[0] = t.x, v[1] = t.y, v[2] = t.z, v[3] = t.w;
  } else {
    v[0] = __ldg(p);
  }
}
template <int VEC>
__device__ __forceinline__ void store_vec(float* p, const float (&v)[VEC]) {
  if constexpr (VEC == 4) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  } else {
    *p = v[0];
  }
}

// ---------------------------------------------------------------------------------------------
// forward: one row per `TPR` threads (32 = a warp, or the whole 256-thread CTA), ITERS vectors
// of VEC floats per thread, all held in registers.
// ---------------------------------------------------------------------------------------------
template <int TPR, int VEC, int ITERS>
__global__ void __launch_bounds__(256) attn_softmax_fw_kernel(float* __restrict__ inp,
                                                              const float* __restrict__ attn_mask,
                                                              long long rows, int nhead, int from_len,
                                                              int to_len, int mask_future) {
  __shared__ float red[32];
  constexpr int ROWS_PER_CTA = 256 / TPR;
  const int sub = threadIdx.x / TPR;  // which row of this CTA
  const int t = threadIdx.x % TPR;    // thread within the row
  // Software-pipelined row loop: the next row's values are requested before the two dependent
  // reductions of the current row, so HBM latency overlaps the reduction latency.
  const long long stride = static_cast<long long>(gridDim.x) * ROWS_PER_CTA;
  float nv[ITERS][VEC];
  auto fetch = [&](long long r0) {
    const long long r = r0 + sub;
    const float* xr = inp + ((r < rows) ? r : rows - 1) * to_len;
#pragma unroll
    for (int it = 0; it < ITERS; ++it) {
      const int c0 = (it * TPR + t) * VEC;
      if (c0 < to_len) load_vec<VEC>(xr + c0, nv[it]);
    }
  };
  long long row0 = static_cast<long long>(blockIdx.x) * ROWS_PER_CTA;
  if (row0 < rows) fetch(row0);
  for (; row0 < rows; row0 += stride) {
    const long long row = row0 + sub;
    const bool row_ok = row < rows;
    const long long rr = row_ok ? row : rows - 1;  // keep every thread in the collectives
    const int q = static_cast<int>(rr % from_len);
    const long long b = rr / (static_cast<long long>(from_len) * nhead);
    float* x = inp + rr * to_len;
    const float* mk = attn_mask ? attn_mask + b * to_len : nullptr;

    float v[ITERS][VEC];
    float mx = -FLT_MAX;
#pragma unroll
    for (int it = 0; it < ITERS; ++it) {
      const int c0 = (it * TPR + t) * VEC;
      if (c0 < to_len) {
#pragma unroll
        for (int k = 0; k < VEC; ++k) v[it][k] = nv[it][k];
        if (mk) {
          float m4[VEC];
          load_vec<VEC>(mk + c0, m4);
#pragma unroll
          for (int k = 0; k < VEC; ++k) v[it][k] += m4[k];
        }
        if (mask_future) {
#pragma unroll
          for (int k = 0; k < VEC; ++k)
            if (c0 + k > q) v[it][k] = kMaskedScore;
        }
      } else {
#pragma unroll
        for (int k = 0; k < VEC; ++k) v[it][k] = kMaskedScore;
      }
#pragma unroll
      for (int k = 0; k < VEC; ++k) mx = fmaxf(mx, v[it][k]);
    }
    if (row0 + stride < rows) fetch(row0 + stride);
    mx = row_reduce<true, TPR>(mx, red);
    float sum = 0.f;
#pragma unroll
    for (int it = 0; it < ITERS; ++it) {
      const int c0 = (it * TPR + t) * VEC;
#pragma unroll
      for (int k = 0; k < VEC; ++k) {
        // columns >= to_len must not contribute (the reference pads them with -1e8, which only
        // matters if the whole row is masked; keep that behaviour for in-range columns only)
        const float e = (c0 + k < to_len) ? __expf(v[it][k] - mx) : 0.f;
        v[it][k] = e;
        sum += e;
      }
    }
    sum = row_reduce<false, TPR>(sum, red);
    const float inv = __fdividef(1.0f, sum + kSoftmaxEps);
    if (row_ok) {
#pragma unroll
      for (int it = 0; it < ITERS; ++it) {
        const int c0 = (it * TPR + t) * VEC;
        if (c0 < to_len) {
#pragma unroll
          for (int k = 0; k < VEC; ++k) v[it][k] *= inv;
          store_vec<VEC>(x + c0, v[it]);
        }
      }
    }
  }
}

// backward: grad <- y * (grad - sum_j y*grad)
template <int TPR, int VEC, int ITERS>
__global__ void __launch_bounds__(256) attn_softmax_bw_kernel(float* __restrict__ grad,
                                                              const float* __restrict__ y, long long rows,
                                                              int len) {
  __shared__ float red[32];
  constexpr int ROWS_PER_CTA = 256 / TPR;
  const int sub = threadIdx.x / TPR;
  const int t = threadIdx.x % TPR;
  for (long long row0 = static_cast<long long>(blockIdx.x) * ROWS_PER_CTA; row0 < rows;
       row0 += static_cast<long long>(gridDim.x) * ROWS_PER_CTA) {
    const long long row = row0 + sub;
    const bool row_ok = row < rows;
    const long long rr = row_ok ? row : rows - 1;
    float* g = grad + rr * len;
    const float* yy = y + rr * len;
    float gv[ITERS][VEC], yv[ITERS][VEC];
    float s = 0.f;
#pragma unroll
    for (int it = 0; it < ITERS; ++it) {
      const int c0 = (it * TPR + t) * VEC;
      if (c0 < len) {
        load_vec<VEC>(g + c0, gv[it]);
        load_vec<VEC>(yy + c0, yv[it]);
#pragma unroll
        for (int k = 0; k < VEC; ++k) s += gv[it][k] * yv[it][k];
      }
    }
    s = row_reduce<false, TPR>(s, red);
    if (row_ok) {
#pragma unroll
      for (int it = 0; it < ITERS; ++it) {
        const int c0 = (it * TPR + t) * VEC;
        if (c0 < len) {
#pragma unroll
          for (int k = 0; k < VEC; ++k) gv[it][k] = yv[it][k] * (gv[it][k] - s);
          store_vec<VEC>(g + c0, gv[it]);
        }
      }
    }
  }
}

static int num_sms() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

// Pick (TPR, VEC, ITERS) for a row length; returns false if the row is too long.
template <typename F>
static bool dispatch_row(int len, bool aligned, F&& f) {
  // f.template operator()<TPR, VEC, ITERS>()
  if (aligned && len % 4 == 0) {
    const int v = len / 4;  // float4 per row
    if (v <= 32) return f.template operator()<32, 4, 1>(), true;
    if (v <= 64) return f.template operator()<32, 4, 2>(), true;
    if (v <= 128) return f.template operator()<32, 4, 4>(), true;
    if (v <= 256) return f.template operator()<32, 4, 8>(), true;
    if (v <= 512) return f.template operator()<32, 4, 16>(), true;
    if (v <= 1024) return f.template operator()<256, 4, 4>(), true;
    if (v <= 2048) return f.template operator()<256, 4, 8>(), true;
    if (v <= 4096) return f.template operator()<256, 4, 16>(), true;
    return false;
  }
  if (len <= 32) return f.template operator()<32, 1, 1>(), true;
  if (len <= 64) return f.template operator()<32, 1, 2>(), true;
  if (len <= 128) return f.template operator()<32, 1, 4>(), true;
  if (len <= 256) return f.template operator()<32, 1, 8>(), true;
  if (len <= 512) return f.template operator()<32, 1, 16>(), true;
  if (len <= 1024) return f.template operator()<32, 1, 32>(), true;
  if (len <= 4096) return f.template operator()<256, 1, 16>(), true;
  if (len <= 16384) return f.template operator()<256, 1, 64>(), true;
  return false;
}

// Forward dispatch: at most 4 values per thread up to 1024 columns (the two dependent reductions
// make the row latency-bound, so occupancy matters more than work per thread: <32,4,16> at 2048
// columns ran at 155 registers / 12% occupancy / 27% of HBM peak).
template <typename F>
static bool dispatch_row_fw(int len, bool aligned, F&& f) {
  if (aligned && len % 4 == 0) {
    const int v = len / 4;
    // up to 1024 columns a WARP owns a row (shuffle reductions only; the 64/128/256-thread-per-row variants paid
    // four block barriers per pair of rows: 56 % of the HBM peak at 512 columns against 111 % for the warp-per-row
    // backward on the same tensor); above that the row is spread over the CTA to keep the register count down
    if (v <= 32) return f.template operator()<32, 4, 1>(), true;
    if (v <= 64) return f.template operator()<32, 4, 2>(), true;
    if (v <= 128) return f.template operator()<32, 4, 4>(), true;
    if (v <= 256) return f.template operator()<32, 4, 8>(), true;
    if (v <= 512) return f.template operator()<256, 4, 2>(), true;
    if (v <= 1024) return f.template operator()<256, 4, 4>(), true;
    if (v <= 2048) return f.template operator()<256, 4, 8>(), true;
    if (v <= 4096) return f.template operator()<256, 4, 16>(), true;
    return false;
  }
  return dispatch_row(len, false, f);
}

static int grid_for(long long rows, int rows_per_cta) {
  long long need = (rows + rows_per_cta - 1) / rows_per_cta;
  long long cap = static_cast<long long>(num_sms()) * 16;  // 16 resident CTAs/SM worth of waves
  if (need < 1) need = 1;
  return static_cast<int>(need < cap ? need : cap);
}

}  // namespace fa

extern "C" {

int fa_attn_softmax_dev(float* inp, const float* attn_mask, int batch_size, int nhead, int from_len, int to_len,
                        int mask_future, fa_stream_t stream) {
  fa::clear_error();
  if (batch_size < 0 || nhead < 0 || from_len < 0 || to_len <= 0 || !inp)
    return fa::set_error(FA_ERR_INVALID, "attn_softmax: bad shape (%d,%d,%d,%d)", batch_size, nhead, from_len,
                         to_len);
  const long long rows = static_cast<long long>(batch_size) * nhead * from_len;
  if (rows == 0) return FA_OK;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const bool aligned = (reinterpret_cast<uintptr_t>(inp) % 16 == 0) &&
                       (!attn_mask || reinterpret_cast<uintptr_t>(attn_mask) % 16 == 0);
  bool ok = fa::dispatch_row_fw(to_len, aligned, [&]<int TPR, int VEC, int ITERS>() {
    fa::attn_softmax_fw_kernel<TPR, VEC, ITERS><<<fa::grid_for(rows, 256 / TPR), 256, 0, s>>>(
        inp, attn_mask, rows, nhead, from_len, to_len, mask_future);
    fa::count_launch();
  });
  if (!ok) return fa::set_error(FA_ERR_UNSUPPORTED, "attn_softmax: to_len %d > 16384 not supported", to_len);
  FA_CUDA_CHECK(cudaGetLastError());
  return FA_OK;
}

int fa_attn_softmax_bw_dev(float* out_grad, const float* soft_inp, long long rows, int softmax_len,
                           fa_stream_t stream) {
  fa::clear_error();
  if (rows < 0 || softmax_len <= 0 || !out_grad || !soft_inp)
    return fa::set_error(FA_ERR_INVALID, "attn_softmax_bw: bad shape (%lld,%d)", rows, softmax_len);
  if (rows == 0) return FA_OK;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const bool aligned =
      (reinterpret_cast<uintptr_t>(out_grad) % 16 == 0) && (reinterpret_cast<uintptr_t>(soft_inp) % 16 == 0);
  bool ok = fa::dispatch_row(softmax_len, aligned, [&]<int TPR, int VEC, int ITERS>() {
    fa::attn_softmax_bw_kernel<TPR, VEC, ITERS>
        <<<fa::grid_for(rows, 256 / TPR), 256, 0, s>>>(out_grad, soft_inp, rows, softmax_len);
    fa::count_launch();
  });
  if (!ok)
    return fa::set_error(FA_ERR_UNSUPPORTED, "attn_softmax_bw: softmax_len %d > 16384 not supported",
                         softmax_len);
  FA_CUDA_CHECK(cudaGetLastError());
  return FA_OK;
}

// ---- legacy host-pointer ABI (reference src/softmax_kernel.cu:233, :345) ---------------------
void launch_attn_softmax(float* inp, const float* attn_mask, int batch_size, int nhead, int from_len,
                         int to_len, bool mask_future, fa_stream_t stream) {
  fa::clear_error();
  const size_t n = static_cast<size_t>(batch_size) * nhead * from_len * to_len;
  const size_t nm = static_cast<size_t>(batch_size) * to_len;
  if (n == 0) return;
  float* d_inp = static_cast<float*>(fa::g_pool.get(0, n * sizeof(float)));
  float* d_mask = attn_mask ? static_cast<float*>(fa::g_pool.get(1, nm * sizeof(float))) : nullptr;
  if (!d_inp || (attn_mask && !d_mask)) {
    fa::set_error(FA_ERR_CUDA, "launch_attn_softmax: device allocation of %zu bytes failed", n * sizeof(float));
    return;
  }
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  auto fail = [&](cudaError_t e, const char* what) {
    fa::set_error(FA_ERR_CUDA, "launch_attn_softmax: %s: %s", what, cudaGetErrorString(e));
  };
  cudaError_t e;
  if ((e = cudaMemcpyAsync(d_inp, inp, n * sizeof(float), cudaMemcpyHostToDevice, s)) != cudaSuccess)
    return fail(e, "H2D inp");
  if (attn_mask &&
      (e = cudaMemcpyAsync(d_mask, attn_mask, nm * sizeof(float), cudaMemcpyHostToDevice, s)) != cudaSuccess)
    return fail(e, "H2D mask");
  if (fa_attn_softmax_dev(d_inp, d_mask, batch_size, nhead, from_len, to_len, mask_future ? 1 : 0, stream) !=
      FA_OK)
    return;
  if ((e = cudaMemcpyAsync(inp, d_inp, n * sizeof(float), cudaMemcpyDeviceToHost, s)) != cudaSuccess)
    return fail(e, "D2H");
  if ((e = cudaStreamSynchronize(s)) != cudaSuccess) return fail(e, "sync");
}

void launch_attn_softmax_bw(float* out_grad, const float* soft_inp, int rows, int softmax_len,
                            fa_stream_t stream) {
  fa::clear_error();
  const size_t n = static_cast<size_t>(rows) * softmax_len;
  if (n == 0) return;
  float* d_g = static_cast<float*>(fa::g_pool.get(0, n * sizeof(float)));
  float* d_y = static_cast<float*>(fa::g_pool.get(1, n * sizeof(float)));
  if (!d_g || !d_y) {
    fa::set_error(FA_ERR_CUDA, "launch_attn_softmax_bw: device allocation of %zu bytes failed",
                  n * sizeof(float));
    return;
  }
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  auto fail = [&](cudaError_t e, const char* what) {
    fa::set_error(FA_ERR_CUDA, "launch_attn_softmax_bw: %s: %s", what, cudaGetErrorString(e));
  };
  cudaError_t e;
  if ((e = cudaMemcpyAsync(d_g, out_grad, n * sizeof(float), cudaMemcpyHostToDevice, s)) != cudaSuccess)
    return fail(e, "H2D grad");
  if ((e = cudaMemcpyAsync(d_y, soft_inp, n * sizeof(float), cudaMemcpyHostToDevice, s)) != cudaSuccess)
    return fail(e, "H2D y");
  if (fa_attn_softmax_bw_dev(d_g, d_y, rows, softmax_len, stream) != FA_OK) return;
  if ((e = cudaMemcpyAsync(out_grad, d_g, n * sizeof(float), cudaMemcpyDeviceToHost, s)) != cudaSuccess)
    return fail(e, "D2H");
  if ((e = cudaStreamSynchronize(s)) != cudaSuccess) return fail(e, "sync");
}

}  // extern "C"
