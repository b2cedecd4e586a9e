// layernorm_kernel.so -- LayerNorm forward / backward for sm_100a.
//
// Replaces the reference's LightSeq-derived src/layernorm_kernel.cu (ker_layer_norm :37,
// ker_ln_bw_dgamma_dbetta :193, ker_ln_bw_dinp :292) behind the same C ABI (launch_layernorm
// :101, launch_layernorm_bw :370).  Both directions are HBM-bound; the design goal is the
// algorithmic minimum traffic with 128-bit coalesced accesses:
//   forward : x read once (registers), y written once            -> 8 B/elem
//   backward: x and dy read once, dx written once; dgamma/dbeta are accumulated in registers
//             across the rows a CTA owns and reduced through a tiny [grid][h] workspace
//             -> 12 B/elem (the reference reads x and dy twice: 20 B/elem)
// Wide rows (2048..4096) in the backward are staged through shared memory with cp.async.bulk.
// Semantics kept from the reference: `vars` stores var + 1e-8 (:70) and the backward adds
// 1e-8 again (:229, :310); hidden_dim % 4 == 0 is required (:105); the 4096 cap of the
// backward (:411) is lifted to 16384.
#include <cstdint>

#include "host_common.cuh"

namespace fa {

constexpr float kLnEps = 1e-8f;  // LN_EPSILON, reference src/layernorm_kernel.cu:12

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Sum of `v` over the TPR threads that own one row.  TPR == 32: a warp shuffle.  Otherwise the
// row is owned by the whole CTA (blockDim.x == TPR).
template <int TPR>
__device__ __forceinline__ float row_sum(float v, float* red) {
  v = warp_sum(v);
  if constexpr (TPR == 32) {
    return v;
  } else {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    constexpr int NW = TPR / 32;
    __syncthreads();
    if (lane == 0) red[w] = v;
    __syncthreads();
    float r = (lane < NW) ? red[lane] : 0.f;
    return warp_sum(r);
  }
}

// ---------------------------------------------------------------------------------------------
// forward.  BLOCK threads; a row is owned by TPR threads, each holding ITERS float4.
// ---------------------------------------------------------------------------------------------
template <int BLOCK, int TPR, int ITERS>
__global__ void __launch_bounds__(BLOCK) layernorm_fw_kernel(float* __restrict__ y, float* __restrict__ vars,
                                                             float* __restrict__ means,
                                                             const float* __restrict__ x,
                                                             const float* __restrict__ gamma,
                                                             const float* __restrict__ beta, long long rows,
                                                             int h) {
  __shared__ float red[32];
  constexpr int RPC = BLOCK / TPR;
  const int sub = threadIdx.x / TPR, t = threadIdx.x % TPR;
  const int h4 = h >> 2;
  const float inv_h = 1.0f / static_cast<float>(h);
  for (long long row0 = static_cast<long long>(blockIdx.x) * RPC; row0 < rows;
       row0 += static_cast<long long>(gridDim.x) * RPC) {
    const long long row = row0 + sub;
    const bool ok = row < rows;
    const long long rr = ok ? row : rows - 1;
    const float4* xr = reinterpret_cast<const float4*>(x + rr * h);
    float4 v[ITERS];
    float s = 0.f;
#pragma unroll
    for (int it = 0; it < ITERS; ++it) {
      const int c = it * TPR + t;
      v[it] = (c < h4) ? __ldg(xr + c) : make_float4(0.f, 0.f, 0.f, 0.f);
      s += (v[it].x + v[it].y) + (v[it].z + v[it].w);
    }
    const float mean = row_sum<TPR>(s, red) * inv_h;
    float sq = 0.f;
#pragma unroll
    for (int it = 0; it < ITERS; ++it) {
      const int c = it * TPR + t;
      if (c < h4) {
        const float a = v[it].x - mean, b = v[it].y - mean, cc = v[it].z - mean, d = v[it].w - mean;
        sq += (a * a + b * b) + (cc * cc + d * d);
      }
    }
    const float var = row_sum<TPR>(sq, red) * inv_h + kLnEps;  // stored WITH eps (reference :70)
    const float rstd = rsqrtf(var);
    if (ok) {
      if (t == 0) {
        means[row] = mean;
        vars[row] = var;
      }
      float4* yr = reinterpret_cast<float4*>(y + row * h);
#pragma unroll
      for (int it = 0; it < ITERS; ++it) {
        const int c = it * TPR + t;
        if (c < h4) {
          const float4 g = __ldg(reinterpret_cast<const float4*>(gamma) + c);
          const float4 b = __ldg(reinterpret_cast<const float4*>(beta) + c);
          float4 o;
          o.x = g.x * ((v[it].x - mean) * rstd) + b.x;
          o.y = g.y * ((v[it].y - mean) * rstd) + b.y;
          o.z = g.z * ((v[it].z - mean) * rstd) + b.z;
          o.w = g.w * ((v[it].w - mean) * rstd) + b.w;
          yr[c] = o;
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// backward, fused: dx per row + per-CTA partial dgamma / dbeta.
//   xhat = (x - mean) / sqrt(var + eps)           (reference :229, :310 -- eps a second time)
//   dxhat = dy * gamma
//   dx = (dxhat - (sum(dxhat) + xhat * sum(dxhat*xhat)) / h) / sd      (reference :348-362)
// part_g / part_b: [gridDim.x][h] fp32 partial sums, reduced by ln_bw_reduce_kernel.
// ---------------------------------------------------------------------------------------------
template <int BLOCK, int TPR, int ITERS>
__global__ void __launch_bounds__(BLOCK) layernorm_bw_kernel(float* __restrict__ dx, float* __restrict__ part_g,
                                                             float* __restrict__ part_b,
                                                             const float* __restrict__ dy,
                                                             const float* __restrict__ x,
                                                             const float* __restrict__ gamma,
                                                             const float* __restrict__ vars,
                                                             const float* __restrict__ means, long long rows,
                                                             int h) {
  extern __shared__ float4 sm_part[];  // RPC > 1: [2][RPC][h4] cross-row-group reduction buffer
  __shared__ float red[32];
  constexpr int RPC = BLOCK / TPR;
  const int sub = threadIdx.x / TPR, t = threadIdx.x % TPR;
  const int h4 = h >> 2;
  const float inv_h = 1.0f / static_cast<float>(h);

  float4 g4[ITERS], acc_g[ITERS], acc_b[ITERS];
#pragma unroll
  for (int it = 0; it < ITERS; ++it) {
    const int c = it * TPR + t;
    g4[it] = (c < h4) ? __ldg(reinterpret_cast<const float4*>(gamma) + c) : make_float4(0.f, 0.f, 0.f, 0.f);
    acc_g[it] = make_float4(0.f, 0.f, 0.f, 0.f);
    acc_b[it] = make_float4(0.f, 0.f, 0.f, 0.f);
  }

  for (long long row0 = static_cast<long long>(blockIdx.x) * RPC; row0 < rows;
       row0 += static_cast<long long>(gridDim.x) * RPC) {
    const long long row = row0 + sub;
    const bool ok = row < rows;
    const long long rr = ok ? row : rows - 1;
    const float4* xr = reinterpret_cast<const float4*>(x + rr * h);
    const float4* dyr = reinterpret_cast<const float4*>(dy + rr * h);
    const float mean = __ldg(means + rr);
    const float rstd = rsqrtf(__ldg(vars + rr) + kLnEps);
    float4 xh[ITERS], dxh[ITERS];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int it = 0; it < ITERS; ++it) {
      const int c = it * TPR + t;
      if (c < h4) {
        const float4 xv = __ldg(xr + c);
        const float4 dv = __ldg(dyr + c);
        xh[it] = make_float4((xv.x - mean) * rstd, (xv.y - mean) * rstd, (xv.z - mean) * rstd,
                             (xv.w - mean) * rstd);
        dxh[it] = make_float4(dv.x * g4[it].x, dv.y * g4[it].y, dv.z * g4[it].z, dv.w * g4[it].w);
        if (ok) {
          acc_b[it].x += dv.x, acc_b[it].y += dv.y, acc_b[it].z += dv.z, acc_b[it].w += dv.w;
          acc_g[it].x += dv.x * xh[it].x, acc_g[it].y += dv.y * xh[it].y;
          acc_g[it].z += dv.z * xh[it].z, acc_g[it].w += dv.w * xh[it].w;
        }
        s1 += (dxh[it].x + dxh[it].y) + (dxh[it].z + dxh[it].w);
        s2 += (dxh[it].x * xh[it].x + dxh[it].y * xh[it].y) + (dxh[it].z * xh[it].z + dxh[it].w * xh[it].w);
      } else {
        xh[it] = make_float4(0.f, 0.f, 0.f, 0.f);
        dxh[it] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
    s1 = row_sum<TPR>(s1, red) * inv_h;
    s2 = row_sum<TPR>(s2, red) * inv_h;
    if (ok) {
      float4* dxr = reinterpret_cast<float4*>(dx + row * h);
#pragma unroll
      for (int it = 0; it < ITERS; ++it) {
        const int c = it * TPR + t;
        if (c < h4) {
          float4 o;
          o.x = (dxh[it].x - s1 - xh[it].x * s2) * rstd;
          o.y = (dxh[it].y - s1 - xh[it].y * s2) * rstd;
          o.z = (dxh[it].z - s1 - xh[it].z * s2) * rstd;
          o.w = (dxh[it].w - s1 - xh[it].w * s2) * rstd;
          dxr[c] = o;
        }
      }
    }
  }

  // fold the RPC row groups of this CTA together, then publish one partial row per CTA
  float4* pg = reinterpret_cast<float4*>(part_g + static_cast<size_t>(blockIdx.x) * h);
  float4* pb = reinterpret_cast<float4*>(part_b + static_cast<size_t>(blockIdx.x) * h);
  if constexpr (RPC == 1) {
#pragma unroll
    for (int it = 0; it < ITERS; ++it) {
      const int c = it * TPR + t;
      if (c < h4) pg[c] = acc_g[it], pb[c] = acc_b[it];
    }
  } else {
    float4* sg = sm_part;                 // [RPC][h4]
    float4* sb = sm_part + RPC * h4;      // [RPC][h4]
#pragma unroll
    for (int it = 0; it < ITERS; ++it) {
      const int c = it * TPR + t;
      if (c < h4) sg[sub * h4 + c] = acc_g[it], sb[sub * h4 + c] = acc_b[it];
    }
    __syncthreads();
    for (int c = threadIdx.x; c < h4; c += BLOCK) {
      float4 a = sg[c], b = sb[c];
#pragma unroll
      for (int r = 1; r < RPC; ++r) {
        const float4 a2 = sg[r * h4 + c], b2 = sb[r * h4 + c];
        a.x += a2.x, a.y += a2.y, a.z += a2.z, a.w += a2.w;
        b.x += b2.x, b.y += b2.y, b.z += b2.z, b.w += b2.w;
      }
      pg[c] = a, pb[c] = b;
    }
  }
}

// dgamma[c] = sum_p part_g[p][c].  32x32 threads per CTA: threadIdx.x walks 32 consecutive columns
// (coalesced 128-byte rows of the partial matrix), threadIdx.y strides over the partial rows, then a
// shared-memory transpose + warp shuffle folds the 32 row-strides.  (One thread per column with a
// serial loop over ~300 partial rows was latency-bound and cost as much as the main kernel.)
__global__ void __launch_bounds__(1024) ln_bw_reduce_kernel(float* __restrict__ dgamma, float* __restrict__ dbeta,
                                                            const float* __restrict__ part_g,
                                                            const float* __restrict__ part_b, int nparts, int h) {
  __shared__ float sg[32][33], sb[32][33];
  const int c = blockIdx.x * 32 + threadIdx.x;
  float g = 0.f, b = 0.f;
  if (c < h) {
    for (int p = threadIdx.y; p < nparts; p += 32) {
      g += __ldg(part_g + static_cast<size_t>(p) * h + c);
      b += __ldg(part_b + static_cast<size_t>(p) * h + c);
    }
  }
  sg[threadIdx.y][threadIdx.x] = g;
  sb[threadIdx.y][threadIdx.x] = b;
  __syncthreads();
  g = sg[threadIdx.x][threadIdx.y];   // transpose: lanes now hold the 32 row-strides of column threadIdx.y
  b = sb[threadIdx.x][threadIdx.y];
  g = warp_sum(g);
  b = warp_sum(b);
  const int co = blockIdx.x * 32 + threadIdx.y;
  if (threadIdx.x == 0 && co < h) {
    dgamma[co] = g;
    dbeta[co] = b;
  }
}

// ---------------------------------------------------------------------------------------------
// backward for wide rows (2048 <= hidden <= 4096): rows are staged through shared memory by bulk
// async copies (cp.async.bulk, the 1-D TMA path) issued by a producer warp several rows ahead, so
// HBM latency never sits in front of the per-row reductions.  Two persistent CTAs per SM: 8 compute
// warps (a row = 256 threads x ITERS float4) + 1 producer warp; a stage = one x row + one dy row.
// The compute threads copy their slice of the stage into registers and release it at once.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t ln_smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void ln_mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok = 0;
  while (!ok) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
        : "=r"(ok)
        : "r"(ln_smem_u32(bar)), "r"(parity)
        : "memory");
  }
}

template <int ITERS>
__global__ void __launch_bounds__(288, 2)
    layernorm_bw_staged_kernel(float* __restrict__ dx, float* __restrict__ part_g, float* __restrict__ part_b,
                               const float* __restrict__ dy, const float* __restrict__ x,
                               const float* __restrict__ gamma, const float* __restrict__ vars,
                               const float* __restrict__ means, long long rows, int h, int nstage) {
  extern __shared__ __align__(128) uint8_t ln_smem[];
  __shared__ uint64_t full[8], empty[8];
  __shared__ float2 red2[8];
  const int h4 = h >> 2;
  const size_t row_bytes = static_cast<size_t>(h) * 4;
  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  const float inv_h = 1.0f / static_cast<float>(h);
  if (t == 0) {
    for (int i = 0; i < nstage; ++i) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(ln_smem_u32(&full[i])), "r"(1) : "memory");
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(ln_smem_u32(&empty[i])), "r"(256) : "memory");
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const long long first = blockIdx.x, stride = gridDim.x;
  const long long nmine = (rows > first) ? (rows - first + stride - 1) / stride : 0;

  if (warp == 8) {
    // ---- producer: one lane streams this CTA's rows into the stage ring
    if (lane == 0) {
      for (long long i = 0; i < nmine; ++i) {
        const int st = static_cast<int>(i % nstage);
        const uint32_t ph = static_cast<uint32_t>((i / nstage) & 1);
        ln_mbar_wait(&empty[st], ph ^ 1);
        const long long row = first + i * stride;
        uint8_t* dst = ln_smem + static_cast<size_t>(st) * 2 * row_bytes;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(ln_smem_u32(&full[st])),
                     "r"(static_cast<uint32_t>(2 * row_bytes))
                     : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                         ln_smem_u32(dst)),
                     "l"(x + row * h), "r"(static_cast<uint32_t>(row_bytes)), "r"(ln_smem_u32(&full[st]))
                     : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                         ln_smem_u32(dst + row_bytes)),
                     "l"(dy + row * h), "r"(static_cast<uint32_t>(row_bytes)), "r"(ln_smem_u32(&full[st]))
                     : "memory");
      }
    }
    return;
  }

  // ---- compute warps
  float4 acc_g[ITERS], acc_b[ITERS], g4[ITERS];
#pragma unroll
  for (int it = 0; it < ITERS; ++it) {
    const int c = it * 256 + t;
    acc_g[it] = make_float4(0.f, 0.f, 0.f, 0.f);
    acc_b[it] = make_float4(0.f, 0.f, 0.f, 0.f);
    g4[it] = (c < h4) ? __ldg(reinterpret_cast<const float4*>(gamma) + c) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  for (long long i = 0; i < nmine; ++i) {
    const int st = static_cast<int>(i % nstage);
    const uint32_t ph = static_cast<uint32_t>((i / nstage) & 1);
    const long long row = first + i * stride;
    const float mean = __ldg(means + row);
    const float rstd = rsqrtf(__ldg(vars + row) + kLnEps);
    ln_mbar_wait(&full[st], ph);
    const float4* sx = reinterpret_cast<const float4*>(ln_smem + static_cast<size_t>(st) * 2 * row_bytes);
    const float4* sdy = reinterpret_cast<const float4*>(ln_smem + static_cast<size_t>(st) * 2 * row_bytes + row_bytes);
    float4 xh[ITERS], dxh[ITERS];
#pragma unroll
    for (int it = 0; it < ITERS; ++it) {
      const int c = it * 256 + t;
      if (c < h4) {
        xh[it] = sx[c];
        dxh[it] = sdy[c];
      } else {
        xh[it] = make_float4(mean, mean, mean, mean);
        dxh[it] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(ln_smem_u32(&empty[st])) : "memory");  // stage is in registers
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int it = 0; it < ITERS; ++it) {
      const float4 dv = dxh[it];
      xh[it] = make_float4((xh[it].x - mean) * rstd, (xh[it].y - mean) * rstd, (xh[it].z - mean) * rstd,
                           (xh[it].w - mean) * rstd);
      acc_b[it].x += dv.x, acc_b[it].y += dv.y, acc_b[it].z += dv.z, acc_b[it].w += dv.w;
      acc_g[it].x += dv.x * xh[it].x, acc_g[it].y += dv.y * xh[it].y;
      acc_g[it].z += dv.z * xh[it].z, acc_g[it].w += dv.w * xh[it].w;
      dxh[it] = make_float4(dv.x * g4[it].x, dv.y * g4[it].y, dv.z * g4[it].z, dv.w * g4[it].w);
      s1 += (dxh[it].x + dxh[it].y) + (dxh[it].z + dxh[it].w);
      s2 += (dxh[it].x * xh[it].x + dxh[it].y * xh[it].y) + (dxh[it].z * xh[it].z + dxh[it].w * xh[it].w);
    }
    // row sums over the 256 compute threads (named barrier 1: the producer warp is not involved)
    s1 = warp_sum(s1);
    s2 = warp_sum(s2);
    asm volatile("bar.sync 1, 256;" ::: "memory");
    if (lane == 0) red2[warp] = make_float2(s1, s2);
    asm volatile("bar.sync 1, 256;" ::: "memory");
    s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) s1 += red2[k].x, s2 += red2[k].y;
    s1 *= inv_h;
    s2 *= inv_h;
    float4* dxr = reinterpret_cast<float4*>(dx + row * h);
#pragma unroll
    for (int it = 0; it < ITERS; ++it) {
      const int c = it * 256 + t;
      if (c < h4) {
        float4 o;
        o.x = (dxh[it].x - s1 - xh[it].x * s2) * rstd;
        o.y = (dxh[it].y - s1 - xh[it].y * s2) * rstd;
        o.z = (dxh[it].z - s1 - xh[it].z * s2) * rstd;
        o.w = (dxh[it].w - s1 - xh[it].w * s2) * rstd;
        dxr[c] = o;
      }
    }
  }
  float4* pg = reinterpret_cast<float4*>(part_g + static_cast<size_t>(blockIdx.x) * h);
  float4* pb = reinterpret_cast<float4*>(part_b + static_cast<size_t>(blockIdx.x) * h);
#pragma unroll
  for (int it = 0; it < ITERS; ++it) {
    const int c = it * 256 + t;
    if (c < h4) pg[c] = acc_g[it], pb[c] = acc_b[it];
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

// (BLOCK, TPR, ITERS) by row length in float4 units.
template <typename F>
static bool dispatch_ln(int h4, F&& f) {
  if (h4 <= 32) return f.template operator()<256, 32, 1>(), true;
  if (h4 <= 64) return f.template operator()<256, 32, 2>(), true;
  if (h4 <= 128) return f.template operator()<256, 32, 4>(), true;
  if (h4 <= 256) return f.template operator()<256, 256, 1>(), true;
  if (h4 <= 512) return f.template operator()<256, 256, 2>(), true;
  if (h4 <= 1024) return f.template operator()<256, 256, 4>(), true;
  if (h4 <= 2048) return f.template operator()<512, 512, 4>(), true;
  if (h4 <= 4096) return f.template operator()<512, 512, 8>(), true;
  return false;
}

// Forward only: a WARP owns a row up to 1024 columns (shuffle reductions; the CTA-per-row variant pays four block
// barriers per 4 KB row: 55 % of the HBM peak at (8192, 1024)).
template <typename F>
static bool dispatch_ln_fw(int h4, F&& f) {
  if (h4 > 128 && h4 <= 256) return f.template operator()<256, 32, 8>(), true;
  return dispatch_ln(h4, f);
}

}  // namespace fa

extern "C" {

int fa_layernorm_dev(float* ln_res, float* vars, float* means, const float* inp, const float* scale,
                     const float* bias, long long rows, int hidden_dim, fa_stream_t stream) {
  fa::clear_error();
  if (hidden_dim <= 0 || hidden_dim % 4 != 0)  // reference :105 throws "violate hidden_dim % 4 = 0"
    return fa::set_error(FA_ERR_INVALID, "layernorm: violate hidden_dim %% 4 = 0 (hidden_dim=%d)", hidden_dim);
  if (rows < 0) return fa::set_error(FA_ERR_INVALID, "layernorm: rows=%lld", rows);
  if (rows == 0) return FA_OK;
  if ((reinterpret_cast<uintptr_t>(ln_res) | reinterpret_cast<uintptr_t>(inp) | reinterpret_cast<uintptr_t>(scale) |
       reinterpret_cast<uintptr_t>(bias)) & 15)
    return fa::set_error(FA_ERR_INVALID, "layernorm: tensors must be 16-byte aligned");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  bool ok = fa::dispatch_ln_fw(hidden_dim / 4, [&]<int BLOCK, int TPR, int ITERS>() {
    constexpr int RPC = BLOCK / TPR;
    long long need = (rows + RPC - 1) / RPC;
    long long cap = static_cast<long long>(fa::num_sms()) * (2048 / BLOCK);
    int grid = static_cast<int>(need < cap ? need : cap);
    fa::layernorm_fw_kernel<BLOCK, TPR, ITERS>
        <<<grid, BLOCK, 0, s>>>(ln_res, vars, means, inp, scale, bias, rows, hidden_dim);
    fa::count_launch();
  });
  if (!ok) return fa::set_error(FA_ERR_UNSUPPORTED, "layernorm: hidden_dim %d > 16384", hidden_dim);
  FA_CUDA_CHECK(cudaGetLastError());
  return FA_OK;
}

int fa_layernorm_bw_dev(float* gamma_grad, float* betta_grad, float* inp_grad, const float* out_grad,
                        const float* inp, const float* gamma, const float* betta, const float* vars,
                        const float* means, long long rows, int hidden_dim, fa_stream_t stream) {
  (void)betta;
  fa::clear_error();
  if (hidden_dim <= 0 || hidden_dim % 4 != 0)
    return fa::set_error(FA_ERR_INVALID, "layernorm_bw: hidden_dim %% 4 != 0 (hidden_dim=%d)", hidden_dim);
  if (rows < 0) return fa::set_error(FA_ERR_INVALID, "layernorm_bw: rows=%lld", rows);
  if ((reinterpret_cast<uintptr_t>(inp_grad) | reinterpret_cast<uintptr_t>(out_grad) |
       reinterpret_cast<uintptr_t>(inp) | reinterpret_cast<uintptr_t>(gamma)) & 15)
    return fa::set_error(FA_ERR_INVALID, "layernorm_bw: tensors must be 16-byte aligned");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  if (rows == 0) {
    FA_CUDA_CHECK(cudaMemsetAsync(gamma_grad, 0, sizeof(float) * hidden_dim, s));
    FA_CUDA_CHECK(cudaMemsetAsync(betta_grad, 0, sizeof(float) * hidden_dim, s));
    return FA_OK;
  }
  int status = FA_OK;
  if (hidden_dim >= 2048 && hidden_dim <= 4096) {
    // wide rows: bulk-async staged kernel, two persistent CTAs per SM (two rows in flight per SM)
    int grid = 2 * fa::num_sms();
    if (grid > rows) grid = static_cast<int>(rows);
    float* part = static_cast<float*>(fa::g_pool.get(8, sizeof(float) * 2 * static_cast<size_t>(grid) * hidden_dim));
    if (!part) return fa::set_error(FA_ERR_CUDA, "layernorm_bw: workspace allocation failed");
    float* part_g = part;
    float* part_b = part + static_cast<size_t>(grid) * hidden_dim;
    const size_t stage_bytes = static_cast<size_t>(hidden_dim) * 8;
    int nstage = static_cast<int>((100 * 1024) / stage_bytes);
    if (nstage > 8) nstage = 8;
    const int smem = static_cast<int>(nstage * stage_bytes);
    const int iters = (hidden_dim / 4 + 255) / 256;
    auto launch = [&](auto kern) -> int {
      FA_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
      kern<<<grid, 288, smem, s>>>(inp_grad, part_g, part_b, out_grad, inp, gamma, vars, means, rows, hidden_dim, nstage);
      fa::count_launch();
      FA_CUDA_CHECK(cudaGetLastError());
      return FA_OK;
    };
    int rc = (iters == 1)   ? launch(fa::layernorm_bw_staged_kernel<1>)
             : (iters == 2) ? launch(fa::layernorm_bw_staged_kernel<2>)
             : (iters == 3) ? launch(fa::layernorm_bw_staged_kernel<3>)
                            : launch(fa::layernorm_bw_staged_kernel<4>);
    if (rc != FA_OK) return rc;
    fa::ln_bw_reduce_kernel<<<(hidden_dim + 31) / 32, dim3(32, 32), 0, s>>>(gamma_grad, betta_grad, part_g, part_b, grid,
                                                                            hidden_dim);
    fa::count_launch();
    FA_CUDA_CHECK(cudaGetLastError());
    return FA_OK;
  }
  bool ok = fa::dispatch_ln(hidden_dim / 4, [&]<int BLOCK, int TPR, int ITERS>() {
    constexpr int RPC = BLOCK / TPR;
    long long need = (rows + RPC - 1) / RPC;
    // persistent: ~2 CTAs per SM keeps the partial workspace (grid*h*8 B) small
    long long cap = static_cast<long long>(fa::num_sms()) * ((BLOCK <= 256) ? 4 : 1);
    int grid = static_cast<int>(need < cap ? need : cap);
    float* part = static_cast<float*>(fa::g_pool.get(8, sizeof(float) * 2 * static_cast<size_t>(grid) * hidden_dim));
    if (!part) {
      status = fa::set_error(FA_ERR_CUDA, "layernorm_bw: workspace allocation failed");
      return;
    }
    float* part_g = part;
    float* part_b = part + static_cast<size_t>(grid) * hidden_dim;
    size_t smem = (RPC > 1) ? sizeof(float) * 2 * RPC * hidden_dim : 0;
    fa::layernorm_bw_kernel<BLOCK, TPR, ITERS><<<grid, BLOCK, smem, s>>>(inp_grad, part_g, part_b, out_grad, inp,
                                                                         gamma, vars, means, rows, hidden_dim);
    fa::count_launch();
    fa::ln_bw_reduce_kernel<<<(hidden_dim + 31) / 32, dim3(32, 32), 0, s>>>(gamma_grad, betta_grad, part_g, part_b, grid,
                                                                     hidden_dim);
    fa::count_launch();
  });
  if (status != FA_OK) return status;
  if (!ok) return fa::set_error(FA_ERR_UNSUPPORTED, "layernorm_bw: hidden_dim %d > 16384", hidden_dim);
  FA_CUDA_CHECK(cudaGetLastError());
  return FA_OK;
}

// ---- legacy host-pointer ABI (reference src/layernorm_kernel.cu:101, :370) -------------------
void launch_layernorm(float* ln_res, float* vars, float* means, const float* inp, const float* scale,
                      const float* bias, int batch_size, int hidden_dim, fa_stream_t stream) {
  fa::clear_error();
  if (hidden_dim <= 0 || hidden_dim % 4 != 0) {
    fa::set_error(FA_ERR_INVALID, "launch_layernorm: violate hidden_dim %% 4 = 0 (hidden_dim=%d)", hidden_dim);
    return;
  }
  const size_t n = static_cast<size_t>(batch_size) * hidden_dim;
  if (n == 0) return;
  float* d_y = static_cast<float*>(fa::g_pool.get(0, n * 4));
  float* d_x = static_cast<float*>(fa::g_pool.get(1, n * 4));
  const size_t bpad = (static_cast<size_t>(batch_size) + 3) & ~size_t(3);  // keep gamma/beta 16-byte aligned
  float* d_small = static_cast<float*>(fa::g_pool.get(2, (2 * bpad + 2 * hidden_dim) * 4));
  if (!d_y || !d_x || !d_small) {
    fa::set_error(FA_ERR_CUDA, "launch_layernorm: device allocation failed");
    return;
  }
  float* d_var = d_small;
  float* d_mean = d_small + bpad;
  float* d_g = d_mean + bpad;
  float* d_b = d_g + hidden_dim;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  cudaError_t e = cudaSuccess;
  auto step = [&](cudaError_t r) {
    if (e == cudaSuccess) e = r;
  };
  step(cudaMemcpyAsync(d_x, inp, n * 4, cudaMemcpyHostToDevice, s));
  step(cudaMemcpyAsync(d_g, scale, hidden_dim * 4, cudaMemcpyHostToDevice, s));
  step(cudaMemcpyAsync(d_b, bias, hidden_dim * 4, cudaMemcpyHostToDevice, s));
  if (e == cudaSuccess && fa_layernorm_dev(d_y, d_var, d_mean, d_x, d_g, d_b, batch_size, hidden_dim, stream) != FA_OK)
    return;
  step(cudaMemcpyAsync(ln_res, d_y, n * 4, cudaMemcpyDeviceToHost, s));
  step(cudaMemcpyAsync(vars, d_var, static_cast<size_t>(batch_size) * 4, cudaMemcpyDeviceToHost, s));
  step(cudaMemcpyAsync(means, d_mean, static_cast<size_t>(batch_size) * 4, cudaMemcpyDeviceToHost, s));
  step(cudaStreamSynchronize(s));
  if (e != cudaSuccess) fa::set_error(FA_ERR_CUDA, "launch_layernorm: %s", cudaGetErrorString(e));
}

void launch_layernorm_bw(float* gamma_grad, float* betta_grad, float* inp_grad, const float* out_grad,
                         const float* inp, const float* gamma, const float* betta, const float* vars,
                         const float* means, int batch_size, int hidden_dim, fa_stream_t stream_1,
                         fa_stream_t stream_2) {
  (void)stream_2;  // the fused kernel needs one stream; the reference used two (:404-417)
  fa::clear_error();
  if (hidden_dim <= 0 || hidden_dim % 4 != 0) {
    fa::set_error(FA_ERR_INVALID, "launch_layernorm_bw: hidden_dim %% 4 != 0 (hidden_dim=%d)", hidden_dim);
    return;
  }
  const size_t n = static_cast<size_t>(batch_size) * hidden_dim;
  float* d_dx = static_cast<float*>(fa::g_pool.get(0, n * 4));
  float* d_dy = static_cast<float*>(fa::g_pool.get(1, n * 4));
  float* d_x = static_cast<float*>(fa::g_pool.get(3, n * 4));
  const size_t bpad = (static_cast<size_t>(batch_size) + 3) & ~size_t(3);  // keep gamma/beta 16-byte aligned
  float* d_small = static_cast<float*>(fa::g_pool.get(2, (2 * bpad + 4 * hidden_dim) * 4));
  if (!d_dx || !d_dy || !d_x || !d_small) {
    fa::set_error(FA_ERR_CUDA, "launch_layernorm_bw: device allocation failed");
    return;
  }
  float* d_var = d_small;
  float* d_mean = d_small + bpad;
  float* d_g = d_mean + bpad;
  float* d_b = d_g + hidden_dim;
  float* d_dg = d_b + hidden_dim;
  float* d_db = d_dg + hidden_dim;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream_1);
  cudaError_t e = cudaSuccess;
  auto step = [&](cudaError_t r) {
    if (e == cudaSuccess) e = r;
  };
  step(cudaMemcpyAsync(d_dy, out_grad, n * 4, cudaMemcpyHostToDevice, s));
  step(cudaMemcpyAsync(d_x, inp, n * 4, cudaMemcpyHostToDevice, s));
  step(cudaMemcpyAsync(d_g, gamma, hidden_dim * 4, cudaMemcpyHostToDevice, s));
  if (betta) step(cudaMemcpyAsync(d_b, betta, hidden_dim * 4, cudaMemcpyHostToDevice, s));
  step(cudaMemcpyAsync(d_var, vars, static_cast<size_t>(batch_size) * 4, cudaMemcpyHostToDevice, s));
  step(cudaMemcpyAsync(d_mean, means, static_cast<size_t>(batch_size) * 4, cudaMemcpyHostToDevice, s));
  if (e == cudaSuccess && fa_layernorm_bw_dev(d_dg, d_db, d_dx, d_dy, d_x, d_g, d_b, d_var, d_mean, batch_size,
                                              hidden_dim, stream_1) != FA_OK)
    return;
  step(cudaMemcpyAsync(gamma_grad, d_dg, hidden_dim * 4, cudaMemcpyDeviceToHost, s));
  step(cudaMemcpyAsync(betta_grad, d_db, hidden_dim * 4, cudaMemcpyDeviceToHost, s));
  step(cudaMemcpyAsync(inp_grad, d_dx, n * 4, cudaMemcpyDeviceToHost, s));
  step(cudaStreamSynchronize(s));
  if (e != cudaSuccess) fa::set_error(FA_ERR_CUDA, "launch_layernorm_bw: %s", cudaGetErrorString(e));
}

}  // extern "C"
