// fp32-accurate flash attention (forward, backward) on CUDA cores -- the FA_MODE_FP32 path.
//
// Why it exists: the reference's tests demand 1e-5 agreement with an fp32 oracle and head
// dims from 4 to 1024 (tests/test_flash_attention.py:103-108, :162-179); bf16 tensor-core
// arithmetic cannot deliver that, so this mode keeps every product and sum in fp32 (accurate
// expf) for ANY N and d.  It is a new design, not the reference's one-thread-per-row kernel
// (src/flashattention_kernel.cu:9-112): 64x64 score tiles, a 4x4 register micro-tile per
// thread, row statistics by 16-lane __shfl_xor reductions, O/l/m kept on chip for the whole
// KV sweep (the reference round-trips them through HBM every KV tile, :92-104), causal tiles
// above the diagonal skipped, ragged N handled by predication.
//   forward : grid (q_tiles * d_slices, H, B); O columns beyond 256 go to another d-slice
//   backward: two deterministic kernels (no atomics): dK/dV per KV tile, dQ per Q tile,
//             both recompute P from (m, l) like the reference's backward (:194).
#pragma once
#include <cfloat>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

namespace fa {

struct AttnParams {
  int B, H, N, d;
  int causal;
  long long sb, sh, sn;  // element strides of the 4-D tensors
  const int* kv_len;     // device int32[B] or nullptr
  const float* key_mask; // device (B,N) additive or nullptr
  float scale;           // 1/sqrt(d)
};

namespace f32k {

constexpr int BR = 64, BC = 64, DKC = 32, NT = 256;
constexpr int LDA = DKC + 4;  // 36: float4-aligned, conflict-free for the 8-lane LDS.128 phases
constexpr int LDP = BC + 4;   // 68
constexpr int LDX = 64;

__device__ __forceinline__ float ldf(const float* p) { return __ldg(p); }
__device__ __forceinline__ float ldf(const __nv_bfloat16* p) { return __bfloat162float(*p); }
__device__ __forceinline__ void stf(float* p, float v) { *p = v; }
__device__ __forceinline__ void stf(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }

// smem[r][c] (ld) <- g[(row0+r)*sn + col0 + c] for r<64, c<COLS; zero outside [0,nrows)x[0,ncols)
template <int COLS, int LD, typename T>
__device__ __forceinline__ void load_tile(float* smem, const T* g, long long sn, int row0, int nrows, int col0,
                                          int ncols) {
#pragma unroll
  for (int t = 0; t < (64 * COLS) / NT; ++t) {
    const int idx = threadIdx.x + t * NT;
    const int r = idx / COLS, c = idx % COLS;
    const int gr = row0 + r, gc = col0 + c;
    float v = 0.f;
    if (gr < nrows && gc < ncols) v = ldf(g + static_cast<long long>(gr) * sn + gc);
    smem[r * LD + c] = v;
  }
}

// acc[i][j] += sum_k A[ty*4+i][k] * B[tx+16j][k], k < DKC   (A, B: [64][LDA])
__device__ __forceinline__ void gemm_nt(float (&acc)[4][4], const float* A, const float* Bm, int ty, int tx) {
#pragma unroll
  for (int k = 0; k < DKC; k += 4) {
    float4 a[4], b[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) a[i] = *reinterpret_cast<const float4*>(A + (ty * 4 + i) * LDA + k);
#pragma unroll
    for (int j = 0; j < 4; ++j) b[j] = *reinterpret_cast<const float4*>(Bm + (tx + 16 * j) * LDA + k);
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        acc[i][j] = fmaf(a[i].x, b[j].x, acc[i][j]);
        acc[i][j] = fmaf(a[i].y, b[j].y, acc[i][j]);
        acc[i][j] = fmaf(a[i].z, b[j].z, acc[i][j]);
        acc[i][j] = fmaf(a[i].w, b[j].w, acc[i][j]);
      }
  }
}
// acc[i][j] += sum_k P[ty*4+i][k] * X[k][tx+16j], k < 64   (P: [64][LDP], X: [64][LDX])
__device__ __forceinline__ void gemm_nn(float (&acc)[4][4], const float* P, const float* X, int ty, int tx) {
#pragma unroll 4
  for (int k = 0; k < 64; k += 4) {
    float4 a[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) a[i] = *reinterpret_cast<const float4*>(P + (ty * 4 + i) * LDP + k);
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      float b[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = X[(k + kk) * LDX + tx + 16 * j];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float av = kk == 0 ? a[i].x : kk == 1 ? a[i].y : kk == 2 ? a[i].z : a[i].w;
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av, b[j], acc[i][j]);
      }
    }
  }
}
// acc[i][j] += sum_q P[q][ty*4+i] * X[q][tx+16j], q < 64   (transposed use of P)
__device__ __forceinline__ void gemm_tn(float (&acc)[4][4], const float* P, const float* X, int ty, int tx) {
#pragma unroll 8
  for (int q = 0; q < 64; ++q) {
    const float4 a = *reinterpret_cast<const float4*>(P + q * LDP + ty * 4);
    float b[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) b[j] = X[q * LDX + tx + 16 * j];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      acc[0][j] = fmaf(a.x, b[j], acc[0][j]);
      acc[1][j] = fmaf(a.y, b[j], acc[1][j]);
      acc[2][j] = fmaf(a.z, b[j], acc[2][j]);
      acc[3][j] = fmaf(a.w, b[j], acc[3][j]);
    }
  }
}

// Two-level accumulation for the gradient sums: a tile's 64-term partial sums are formed in a fresh accumulator and
// then added to the running total, so the rounding error of a sum over N keys (queries) grows like sqrt(64) + sqrt(N/64)
// ulps instead of sqrt(N).  (Measured against torch's fp32 CPU GEMMs through the reference's own test at N = 2048:
// the single-chain sums were the largest term of the X.grad error budget.)
__device__ __forceinline__ void add_into(float (&acc)[4][4], const float (&part)[4][4]) {
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] += part[i][j];
}
__device__ __forceinline__ void zero16(float (&part)[4][4]) {
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) part[i][j] = 0.f;
}

// max / sum over the 16 lanes (tx) that share a score row
__device__ __forceinline__ float row16_max(float v) {
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float row16_sum(float v) {
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// 64x64 tile of A.B^T contracted over the full head dim, staged through smem in DKC chunks.
template <typename T>
__device__ __forceinline__ void score_tile(float (&s)[4][4], const T* A, int a_row0, const T* Bm, int b_row0,
                                           const AttnParams& p, float* As, float* Bs, int ty, int tx) {
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) s[i][j] = 0.f;
  for (int dc = 0; dc < p.d; dc += DKC) {
    load_tile<DKC, LDA>(As, A, p.sn, a_row0, p.N, dc, p.d);
    load_tile<DKC, LDA>(Bs, Bm, p.sn, b_row0, p.N, dc, p.d);
    __syncthreads();
    gemm_nt(s, As, Bs, ty, tx);
    __syncthreads();
  }
}

// scaled + masked score for (row, col); -inf when masked out
__device__ __forceinline__ float masked_score(float raw, int row, int col, int kv_end, const float* mrow,
                                              const AttnParams& p) {
  if (col >= kv_end || (p.causal && col > row)) return -INFINITY;
  float v = raw * p.scale;
  if (mrow) v += __ldg(mrow + col);
  return v;
}

constexpr size_t kSmemFwd = sizeof(float) * (2 * 64 * LDA + 64 * LDP + 64 * LDX);
constexpr size_t kSmemBwd = sizeof(float) * (2 * 64 * LDA + 2 * 64 * LDP + 64 * LDX);

// ---------------------------------------------------------------------------------------------
template <typename T, int NCH>
__global__ void __launch_bounds__(NT) fwd_kernel(AttnParams p, const T* __restrict__ Q, const T* __restrict__ K,
                                                 const T* __restrict__ V, T* __restrict__ O,
                                                 float* __restrict__ M, float* __restrict__ L) {
  extern __shared__ __align__(16) float smem[];
  float* As = smem;
  float* Bs = As + 64 * LDA;
  float* Ps = Bs + 64 * LDA;
  float* Xs = Ps + 64 * LDP;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int nqt = (p.N + BR - 1) / BR;
  const int qt = blockIdx.x % nqt, slice = blockIdx.x / nqt;
  const int h = blockIdx.y, b = blockIdx.z;
  const int q0 = qt * BR, col_base = slice * NCH * 64;
  const long long base = b * p.sb + h * p.sh;
  const T* Qb = Q + base;
  const T* Kb = K + base;
  const T* Vb = V + base;
  int kv_end = p.N;
  if (p.kv_len) kv_end = min(kv_end, max(__ldg(p.kv_len + b), 0));
  const float* mrow = p.key_mask ? p.key_mask + static_cast<long long>(b) * p.N : nullptr;
  int nkt = (kv_end + BC - 1) / BC;
  if (p.causal) nkt = min(nkt, qt + 1);

  float m_i[4], l_i[4], acc[NCH][4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i) m_i[i] = -INFINITY, l_i[i] = 0.f;
#pragma unroll
  for (int c = 0; c < NCH; ++c)
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[c][i][j] = 0.f;

  for (int kt = 0; kt < nkt; ++kt) {
    const int k0 = kt * BC;
    float s[4][4];
    score_tile(s, Qb, q0, Kb, k0, p, As, Bs, ty, tx);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int row = q0 + ty * 4 + i;
      float mx = -INFINITY;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        s[i][j] = masked_score(s[i][j], row, k0 + tx + 16 * j, kv_end, mrow, p);
        mx = fmaxf(mx, s[i][j]);
      }
      mx = row16_max(mx);
      const float m_new = fmaxf(m_i[i], mx);
      const float m_safe = (m_new == -INFINITY) ? 0.f : m_new;
      float rs = 0.f;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float e = expf(s[i][j] - m_safe);
        rs += e;
        Ps[(ty * 4 + i) * LDP + tx + 16 * j] = e;
      }
      rs = row16_sum(rs);
      const float alpha = expf(m_i[i] - m_safe);  // exp(-inf) = 0 on the first tile
      l_i[i] = l_i[i] * alpha + rs;
      m_i[i] = m_new;
#pragma unroll
      for (int c = 0; c < NCH; ++c)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[c][i][j] *= alpha;
    }
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
      const int c0 = col_base + c * 64;
      if (c0 < p.d) {
        load_tile<64, LDX>(Xs, Vb, p.sn, k0, p.N, c0, p.d);
        __syncthreads();  // also orders the Ps writes above
        gemm_nn(acc[c], Ps, Xs, ty, tx);
        __syncthreads();
      }
    }
  }

  T* Ob = O + base;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int row = q0 + ty * 4 + i;
    if (row >= p.N) continue;
    const float inv = l_i[i] > 0.f ? 1.0f / l_i[i] : 0.f;
#pragma unroll
    for (int c = 0; c < NCH; ++c)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int col = col_base + c * 64 + tx + 16 * j;
        if (col < p.d) stf(Ob + static_cast<long long>(row) * p.sn + col, acc[c][i][j] * inv);
      }
    if (slice == 0 && tx == 0) {
      const long long r = (static_cast<long long>(b) * p.H + h) * p.N + row;
      M[r] = m_i[i];
      L[r] = l_i[i];
    }
  }
}

// D[r] = sum_x dO*O ; LSE[r] = m + log(l)      (one warp per row)
template <typename T>
__global__ void bwd_prep_kernel(AttnParams p, const T* __restrict__ O, const T* __restrict__ dO,
                                const float* __restrict__ M, const float* __restrict__ L,
                                float* __restrict__ Dv, float* __restrict__ LSE) {
  const long long rows = static_cast<long long>(p.B) * p.H * p.N;
  const int lane = threadIdx.x & 31;
  for (long long r = static_cast<long long>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5); r < rows;
       r += static_cast<long long>(gridDim.x) * (blockDim.x >> 5)) {
    const int n = static_cast<int>(r % p.N);
    const long long bh = r / p.N;
    const long long off = (bh / p.H) * p.sb + (bh % p.H) * p.sh + static_cast<long long>(n) * p.sn;
    float s = 0.f;
    for (int x = lane; x < p.d; x += 32) s += ldf(O + off + x) * ldf(dO + off + x);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) {
      Dv[r] = s;
      const float l = L[r];
      LSE[r] = (l > 0.f) ? M[r] + logf(l) : INFINITY;  // fully masked row -> P = exp(-inf) = 0
    }
  }
}

// Recompute P and dS for one (q tile, kv tile) pair; P -> Ps[q][k], dS -> Ts[q][k].
template <typename T, bool WRITE_P>
__device__ __forceinline__ void recompute_p_ds(const AttnParams& p, const T* Qb, const T* Kb, const T* Vb,
                                               const T* dOb, const float* Dv, const float* LSE, long long rowbase,
                                               int q0, int k0, int kv_end, const float* mrow, float* As, float* Bs,
                                               float* Ps, float* Ts, int ty, int tx) {
  float s[4][4], dp[4][4];
  score_tile(s, Qb, q0, Kb, k0, p, As, Bs, ty, tx);
  score_tile(dp, dOb, q0, Vb, k0, p, As, Bs, ty, tx);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int row = q0 + ty * 4 + i;
    float lse = INFINITY, dv = 0.f;
    if (row < p.N) {
      lse = __ldg(LSE + rowbase + row);
      dv = __ldg(Dv + rowbase + row);
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float sc = masked_score(s[i][j], row, k0 + tx + 16 * j, kv_end, mrow, p);
      const float pr = (row < p.N) ? expf(sc - lse) : 0.f;
      if (WRITE_P) Ps[(ty * 4 + i) * LDP + tx + 16 * j] = pr;
      Ts[(ty * 4 + i) * LDP + tx + 16 * j] = pr * (dp[i][j] - dv);
    }
  }
}

template <typename T, int NCH>
__global__ void __launch_bounds__(NT) bwd_dkdv_kernel(AttnParams p, const T* __restrict__ Q,
                                                      const T* __restrict__ K, const T* __restrict__ V,
                                                      const T* __restrict__ dO, const float* __restrict__ Dv,
                                                      const float* __restrict__ LSE, T* __restrict__ dK,
                                                      T* __restrict__ dV) {
  extern __shared__ __align__(16) float smem[];
  float* As = smem;
  float* Bs = As + 64 * LDA;
  float* Ps = Bs + 64 * LDA;
  float* Ts = Ps + 64 * LDP;
  float* Xs = Ts + 64 * LDP;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int nt = (p.N + 63) / 64;
  const int kt = blockIdx.x % nt, slice = blockIdx.x / nt;
  const int h = blockIdx.y, b = blockIdx.z;
  const int k0 = kt * 64, col_base = slice * NCH * 64;
  const long long base = b * p.sb + h * p.sh;
  const long long rowbase = (static_cast<long long>(b) * p.H + h) * p.N;
  const T *Qb = Q + base, *Kb = K + base, *Vb = V + base, *dOb = dO + base;
  int kv_end = p.N;
  if (p.kv_len) kv_end = min(kv_end, max(__ldg(p.kv_len + b), 0));
  const float* mrow = p.key_mask ? p.key_mask + static_cast<long long>(b) * p.N : nullptr;

  float acc_k[NCH][4][4], acc_v[NCH][4][4];
#pragma unroll
  for (int c = 0; c < NCH; ++c)
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc_k[c][i][j] = 0.f, acc_v[c][i][j] = 0.f;

  if (k0 < kv_end) {
    for (int qt = p.causal ? kt : 0; qt < nt; ++qt) {
      const int q0 = qt * 64;
      recompute_p_ds<T, true>(p, Qb, Kb, Vb, dOb, Dv, LSE, rowbase, q0, k0, kv_end, mrow, As, Bs, Ps, Ts, ty, tx);
#pragma unroll
      for (int c = 0; c < NCH; ++c) {
        const int c0 = col_base + c * 64;
        if (c0 < p.d) {
          float part[4][4];
          load_tile<64, LDX>(Xs, dOb, p.sn, q0, p.N, c0, p.d);
          __syncthreads();
          zero16(part);
          gemm_tn(part, Ps, Xs, ty, tx);
          add_into(acc_v[c], part);
          __syncthreads();
          load_tile<64, LDX>(Xs, Qb, p.sn, q0, p.N, c0, p.d);
          __syncthreads();
          zero16(part);
          gemm_tn(part, Ts, Xs, ty, tx);
          add_into(acc_k[c], part);
          __syncthreads();
        }
      }
    }
  }
  T *dKb = dK + base, *dVb = dV + base;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int row = k0 + ty * 4 + i;
    if (row >= p.N) continue;
#pragma unroll
    for (int c = 0; c < NCH; ++c)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int col = col_base + c * 64 + tx + 16 * j;
        if (col < p.d) {
          stf(dKb + static_cast<long long>(row) * p.sn + col, acc_k[c][i][j] * p.scale);
          stf(dVb + static_cast<long long>(row) * p.sn + col, acc_v[c][i][j]);
        }
      }
  }
}

template <typename T, int NCH>
__global__ void __launch_bounds__(NT) bwd_dq_kernel(AttnParams p, const T* __restrict__ Q,
                                                    const T* __restrict__ K, const T* __restrict__ V,
                                                    const T* __restrict__ dO, const float* __restrict__ Dv,
                                                    const float* __restrict__ LSE, T* __restrict__ dQ) {
  extern __shared__ __align__(16) float smem[];
  float* As = smem;
  float* Bs = As + 64 * LDA;
  float* Ps = Bs + 64 * LDA;
  float* Ts = Ps + 64 * LDP;
  float* Xs = Ts + 64 * LDP;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int nt = (p.N + 63) / 64;
  const int qt = blockIdx.x % nt, slice = blockIdx.x / nt;
  const int h = blockIdx.y, b = blockIdx.z;
  const int q0 = qt * 64, col_base = slice * NCH * 64;
  const long long base = b * p.sb + h * p.sh;
  const long long rowbase = (static_cast<long long>(b) * p.H + h) * p.N;
  const T *Qb = Q + base, *Kb = K + base, *Vb = V + base, *dOb = dO + base;
  int kv_end = p.N;
  if (p.kv_len) kv_end = min(kv_end, max(__ldg(p.kv_len + b), 0));
  const float* mrow = p.key_mask ? p.key_mask + static_cast<long long>(b) * p.N : nullptr;
  int nkt = (kv_end + 63) / 64;
  if (p.causal) nkt = min(nkt, qt + 1);

  float acc[NCH][4][4];
#pragma unroll
  for (int c = 0; c < NCH; ++c)
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[c][i][j] = 0.f;

  for (int kt = 0; kt < nkt; ++kt) {
    const int k0 = kt * 64;
    recompute_p_ds<T, false>(p, Qb, Kb, Vb, dOb, Dv, LSE, rowbase, q0, k0, kv_end, mrow, As, Bs, Ps, Ts, ty, tx);
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
      const int c0 = col_base + c * 64;
      if (c0 < p.d) {
        float part[4][4];
        load_tile<64, LDX>(Xs, Kb, p.sn, k0, p.N, c0, p.d);
        __syncthreads();
        zero16(part);
        gemm_nn(part, Ts, Xs, ty, tx);
        add_into(acc[c], part);
        __syncthreads();
      }
    }
  }
  T* dQb = dQ + base;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int row = q0 + ty * 4 + i;
    if (row >= p.N) continue;
#pragma unroll
    for (int c = 0; c < NCH; ++c)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int col = col_base + c * 64 + tx + 16 * j;
        if (col < p.d) stf(dQb + static_cast<long long>(row) * p.sn + col, acc[c][i][j] * p.scale);
      }
  }
}

}  // namespace f32k
}  // namespace fa
