// Decode-shape attention: ONE query token per (batch, head) against a key/value cache of
// kv_len[b] positions.  The reference has no cache -- project/run_machine_translation.py:299-325 re-runs the whole
// prefix for every generated token -- so this is the kernel behind the cache-aware generate().
//
// HBM-bound: every byte of the visible K and V cache is read exactly once (2 * L * d * sizeof(T) per head), in 16-byte
// vectors, a key row per lane group.  Split-KV: the cache is cut into `nsplit` chunks so that B*H*nsplit CTAs cover the
// 148 SMs even at batch 1; each CTA keeps an online-softmax state (m, l, o) per lane group, merges lane groups by warp
// shuffles and warps through shared memory, and either writes the result (nsplit == 1) or a partial (m, l, o) that a
// small second kernel combines.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cstdint>

namespace fa {
namespace decode {

struct Params {
  int B, H, d;
  int L;                 // number of cached positions when kv_len == nullptr
  const int* kv_len;     // device int32[B] or nullptr
  long long q_sb, q_sh;  // element strides of q and out, logical (B, H, d)
  long long o_sb, o_sh;
  long long c_sb, c_sh, c_sn;   // element strides of the caches, logical (B, H, Lmax, d)
  float scale_log2;      // 1/sqrt(d) * log2(e)
  int nsplit, chunk;     // keys per split
  float* ws_ml;          // (B, H, nsplit, 2)  partial row max (log2 domain) and sum
  float* ws_o;           // (B, H, nsplit, d)  partial un-normalised outputs
  float* lse;            // optional (B, H): natural-log LSE of the scaled scores, or nullptr
};

template <typename T>
struct Vec;
template <>
struct Vec<float> {
  static constexpr int N = 4;
  using Raw = float4;
  __device__ static Raw load_raw(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
  __device__ static void unpack(const Raw& x, float (&v)[4]) { v[0] = x.x, v[1] = x.y, v[2] = x.z, v[3] = x.w; }
  __device__ static void load(const float* p, float (&v)[4]) { unpack(load_raw(p), v); }
};
template <>
struct Vec<__nv_bfloat16> {
  static constexpr int N = 8;
  using Raw = uint4;
  __device__ static Raw load_raw(const __nv_bfloat16* p) { return __ldg(reinterpret_cast<const uint4*>(p)); }
  __device__ static void unpack(const Raw& x, float (&v)[8]) {
    const uint32_t w[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      v[2 * i] = __uint_as_float(w[i] << 16);
      v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
    }
  }
  __device__ static void load(const __nv_bfloat16* p, float (&v)[8]) { unpack(load_raw(p), v); }
};
__device__ __forceinline__ void store_out(float* p, float v) { *p = v; }
__device__ __forceinline__ void store_out(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }

constexpr int kWarps = 4;

// LPK lanes share a key row (power of two <= 32), each lane owns NV 16-byte vectors of it.
template <typename T, int LPK, int NV>
__global__ void __launch_bounds__(kWarps * 32)
    partial_kernel(const Params p, const T* __restrict__ q, const T* __restrict__ kc, const T* __restrict__ vc,
                   T* __restrict__ out) {
  constexpr int VN = Vec<T>::N;
  constexpr int KPW = 32 / LPK;   // keys per warp and step
  const int split = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int sub = lane % LPK, grp = lane / LPK;
  int L = p.kv_len ? __ldg(p.kv_len + b) : p.L;
  L = max(L, 0);
  const int k_begin = split * p.chunk, k_end = min(L, k_begin + p.chunk);
  const int nvec = p.d / VN;      // vectors per row
  // my slice of q
  float qv[NV][VN];
  bool live[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int vi = sub + i * LPK;
    live[i] = vi < nvec;
    if (live[i]) {
      Vec<T>::load(q + b * p.q_sb + h * p.q_sh + vi * VN, qv[i]);
    } else {
#pragma unroll
      for (int x = 0; x < VN; ++x) qv[i][x] = 0.f;
    }
  }
  float m = -INFINITY, l = 0.f;
  float o[NV][VN];
#pragma unroll
  for (int i = 0; i < NV; ++i)
#pragma unroll
    for (int x = 0; x < VN; ++x) o[i][x] = 0.f;
  const T* kb = kc + b * p.c_sb + h * p.c_sh;
  const T* vb = vc + b * p.c_sb + h * p.c_sh;
  // UNR keys per lane group and step: their K and V vectors are all requested before any arithmetic, so a lane keeps
  // 2 * UNR * NV 16-byte loads in flight (with one key per step the kernel ran at 0.55 of the HBM peak: latency-bound).
  // The vectors stay RAW (16 bytes = 4 registers, whatever T is) until they are consumed: unpacking bf16 to floats
  // right after the load cost 115 registers and 4 CTAs per SM, and bf16 caches ran at 0.51 of the HBM peak; raw, the
  // kernel needs 80 (6 CTAs per SM) and reaches it.  (A software-pipelined loop over two register sets was tried on
  // top: 121 registers, 0.65 -- occupancy is worth more here than a warp that never idles.)
  constexpr int UNR = (NV == 1) ? 4 : 2;
  using Raw = typename Vec<T>::Raw;
  const int stride = kWarps * KPW * UNR;
  auto issue = [&](int n0, Raw (&kr)[UNR][NV], Raw (&vr)[UNR][NV]) {
#pragma unroll
    for (int u = 0; u < UNR; ++u) {
      const int n = n0 + u * KPW + grp;
      const long long roff = static_cast<long long>(n < k_end ? n : k_begin) * p.c_sn;   // (clamped: stays in bounds)
#pragma unroll
      for (int i = 0; i < NV; ++i)
        if (live[i]) {
          kr[u][i] = Vec<T>::load_raw(kb + roff + (sub + i * LPK) * VN);
          vr[u][i] = Vec<T>::load_raw(vb + roff + (sub + i * LPK) * VN);
        }
    }
  };
  auto consume = [&](int n0, const Raw (&kr)[UNR][NV], const Raw (&vr)[UNR][NV]) {
    float s[UNR];
    bool ok[UNR];
#pragma unroll
    for (int u = 0; u < UNR; ++u) {
      ok[u] = n0 + u * KPW + grp < k_end;
      s[u] = 0.f;
#pragma unroll
      for (int i = 0; i < NV; ++i)
        if (live[i]) {
          float kf[VN];
          Vec<T>::unpack(kr[u][i], kf);
#pragma unroll
          for (int x = 0; x < VN; ++x) s[u] = fmaf(qv[i][x], kf[x], s[u]);
        }
    }
#pragma unroll
    for (int off = LPK / 2; off > 0; off >>= 1)
#pragma unroll
      for (int u = 0; u < UNR; ++u) s[u] += __shfl_xor_sync(0xffffffffu, s[u], off);
    // online softmax over the (up to) UNR new scores at once
    float m_new = m;
#pragma unroll
    for (int u = 0; u < UNR; ++u)
      if (ok[u]) {
        s[u] *= p.scale_log2;
        m_new = fmaxf(m_new, s[u]);
      }
    if (m_new != -INFINITY) {
      const float corr = exp2f(m - m_new);    // m == -inf on the first key: exp2(-inf) = 0
      float pr[UNR], psum = 0.f;
#pragma unroll
      for (int u = 0; u < UNR; ++u) {
        pr[u] = ok[u] ? exp2f(s[u] - m_new) : 0.f;
        psum += pr[u];
      }
      m = m_new;
      l = l * corr + psum;
#pragma unroll
      for (int i = 0; i < NV; ++i)
        if (live[i]) {
#pragma unroll
          for (int x = 0; x < VN; ++x) o[i][x] *= corr;
#pragma unroll
          for (int u = 0; u < UNR; ++u) {
            float vf[VN];
            Vec<T>::unpack(vr[u][i], vf);
#pragma unroll
            for (int x = 0; x < VN; ++x) o[i][x] = fmaf(pr[u], vf[x], o[i][x]);
          }
        }
    }
  };
  {
    Raw kr[UNR][NV], vr[UNR][NV];
    for (int n0 = k_begin + warp * KPW * UNR; n0 < k_end; n0 += stride) {
      issue(n0, kr, vr);
      consume(n0, kr, vr);
    }
  }
  // merge the lane groups of this warp (lanes with equal `sub` hold the same slice of d)
#pragma unroll
  for (int off = LPK; off < 32; off <<= 1) {
    const float m2 = __shfl_xor_sync(0xffffffffu, m, off), l2 = __shfl_xor_sync(0xffffffffu, l, off);
    const float mn = fmaxf(m, m2);
    const float c1 = (m == -INFINITY) ? 0.f : exp2f(m - mn), c2 = (m2 == -INFINITY) ? 0.f : exp2f(m2 - mn);
    l = l * c1 + l2 * c2;
#pragma unroll
    for (int i = 0; i < NV; ++i)
#pragma unroll
      for (int x = 0; x < VN; ++x) o[i][x] = o[i][x] * c1 + __shfl_xor_sync(0xffffffffu, o[i][x], off) * c2;
    m = mn;
  }
  // merge the warps through shared memory
  __shared__ float s_m[kWarps], s_l[kWarps];
  extern __shared__ float s_o[];   // [kWarps][d]
  if (grp == 0) {
    if (sub == 0) s_m[warp] = m, s_l[warp] = l;
#pragma unroll
    for (int i = 0; i < NV; ++i)
      if (live[i])
#pragma unroll
        for (int x = 0; x < VN; ++x) s_o[warp * p.d + (sub + i * LPK) * VN + x] = o[i][x];
  }
  __syncthreads();
  float mt = -INFINITY;
#pragma unroll
  for (int w = 0; w < kWarps; ++w) mt = fmaxf(mt, s_m[w]);
  float lt = 0.f;
#pragma unroll
  for (int w = 0; w < kWarps; ++w) lt += (s_m[w] == -INFINITY) ? 0.f : s_l[w] * exp2f(s_m[w] - mt);
  const long long unit = (static_cast<long long>(b) * p.H + h);
  for (int x = threadIdx.x; x < p.d; x += blockDim.x) {
    float acc = 0.f;
#pragma unroll
    for (int w = 0; w < kWarps; ++w) acc += (s_m[w] == -INFINITY) ? 0.f : s_o[w * p.d + x] * exp2f(s_m[w] - mt);
    if (p.nsplit == 1)
      store_out(out + b * p.o_sb + h * p.o_sh + x, lt > 0.f ? acc / lt : 0.f);
    else
      p.ws_o[(unit * p.nsplit + split) * p.d + x] = acc;
  }
  if (threadIdx.x == 0) {
    if (p.nsplit == 1) {
      if (p.lse) p.lse[unit] = lt > 0.f ? (mt + log2f(lt)) * 0.6931471805599453f : -INFINITY;
    } else {
      p.ws_ml[(unit * p.nsplit + split) * 2] = mt;
      p.ws_ml[(unit * p.nsplit + split) * 2 + 1] = lt;
    }
  }
}

template <typename T>
__global__ void combine_kernel(const Params p, T* __restrict__ out) {
  const int h = blockIdx.x, b = blockIdx.y;
  const long long unit = (static_cast<long long>(b) * p.H + h);
  const float* ml = p.ws_ml + unit * p.nsplit * 2;
  float mt = -INFINITY;
  for (int s = 0; s < p.nsplit; ++s) mt = fmaxf(mt, ml[2 * s]);
  float lt = 0.f;
  for (int s = 0; s < p.nsplit; ++s) lt += (ml[2 * s] == -INFINITY) ? 0.f : ml[2 * s + 1] * exp2f(ml[2 * s] - mt);
  for (int x = threadIdx.x; x < p.d; x += blockDim.x) {
    float acc = 0.f;
    for (int s = 0; s < p.nsplit; ++s)
      acc += (ml[2 * s] == -INFINITY) ? 0.f : p.ws_o[(unit * p.nsplit + s) * p.d + x] * exp2f(ml[2 * s] - mt);
    store_out(out + b * p.o_sb + h * p.o_sh + x, lt > 0.f ? acc / lt : 0.f);
  }
  if (threadIdx.x == 0 && p.lse) p.lse[unit] = lt > 0.f ? (mt + log2f(lt)) * 0.6931471805599453f : -INFINITY;
}

}  // namespace decode
}  // namespace fa
