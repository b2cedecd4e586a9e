// Flash-attention backward for sm_100a: recompute-based dQ / dK / dV on tcgen05 + TMEM.
//
// One CTA owns one 128-key KV tile of one (batch, head) and sweeps the Q tiles that see it
// (all of them, or only those at / below the diagonal when causal).  K and V stay resident in
// shared memory, the dK and dV accumulators stay resident in TMEM for the whole sweep; per
// Q tile i five tensor-core GEMMs run (all 128x128x128 at d=128):
//     S^T  = K Q_i^T            (SS)  -> TMEM T_S          scores, transposed: lane = key
//     dP^T = V dO_i^T           (SS)  -> TMEM T_dP
//     dV  += P^T dO_i           (TS: P^T is read straight from TMEM, bf16 over T_S)
//     dQ_i = dS K               (SS)  -> TMEM T_dP (reuses the dP columns once dP is consumed)
//     dK  += dS^T Q_i           (SS)
// TMEM: T_S 128 | T_dP/dQ 128 | dV D | dK D columns = 512 at D = 128.
// Working transposed (lane = key) lets P^T feed the dV GEMM from TMEM and puts dS^T into shared
// memory in ONE layout that serves both remaining GEMMs: K-major A for dK, MN-major A for dQ.
// Row statistics LSE_i and D_i = rowsum(dO*O) come from a pre-pass (log2 domain, padded so
// out-of-range queries get P = 0) and arrive per Q tile with the same TMA transaction.
//   warps 0-3 / 4-7 : two compute groups that alternate Q tiles (ping-pong): exp2 -> P^T,
//                     dS^T = P^T*(dP^T - D), then drain dQ_i: TMEM -> registers -> 128B-swizzled
//                     staging box in shared memory -> cp.reduce.async.bulk.tensor (fp32 add) into
//                     the dQ accumulator (LSU red.global tops out near 9 clk per warp instruction
//                     on B200, far too slow for 64 KB per tile pair)
//   warp 8          : TMA producer (K,V once; Q_i + LSE_i + D_i 2-stage ring; dO_i single stage)
//   warp 9          : tcgen05.mma issuer + TMEM allocation
//   warp 10         : issues the dQ TMA reductions and recycles the two staging boxes
// The tensor pipe executes in issue order, so single-buffered T_S / T_dP are enough: S^T(i+1)
// is issued right after dV(i) and overlaps the other group's dS phase.
// dQ is accumulated across KV-tile CTAs in fp32 (TMA add-reductions; summation order varies
// between runs) and converted to bf16 * scale by a small kernel afterwards.
#pragma once
#include "ptx.cuh"

namespace fa {
namespace sm100 {

struct BwdParams {
  int B, H, N, Npad;
  const int* kv_len;
  const float* lse2;   // (B*H, Npad) log2-domain LSE; +inf for rows >= N
  const float* dvec;   // (B*H, Npad) D_i; 0 for rows >= N
  float* dq_acc;       // (B,H,N,D) fp32, zero-initialised
  void* dK;            // bf16 outputs with the strides below
  void* dV;
  long long sb, sh, sn;
  float scale, scale_log2;
  long long* trace;    // bring-up only (FA_TRACE builds): per-iteration clock64() stamps of one CTA
};

#ifdef FA_TRACE
#define FA_TR(slot) \
  if (tr && it < 48) tr[it * 32 + (slot)] = clock64();
#define FA_TRW(slot) \
  if (trw && it < 48) trw[it * 32 + 16 + 4 * w + (slot)] = clock64();
#else
#define FA_TR(slot)
#define FA_TRW(slot)
#endif

template <int D>
struct BwdCfg {
  static constexpr int NCHUNK = D / 64;
  static constexpr int CHUNK_BYTES = 128 * 128;
  static constexpr int TILE_BYTES = NCHUNK * CHUNK_BYTES;   // K, V, Q_i, dO_i tiles: [128][D] bf16
  static constexpr int DS_BYTES = 2 * CHUNK_BYTES;          // dS^T tile: [128 keys][128 queries] bf16
  static constexpr int STG_BYTES = 128 * 128;               // dQ staging box per compute group: [128 q][32 fp32]
  static constexpr int VEC_BYTES = 2 * 2 * 512;             // [stage][lse2 | D][128] fp32
  static constexpr int OFF_K = 0;
  static constexpr int OFF_V = TILE_BYTES;
  static constexpr int OFF_Q = 2 * TILE_BYTES;              // [2 stages]
  static constexpr int OFF_DO = 4 * TILE_BYTES;             // [1 stage]
  static constexpr int OFF_DS = 5 * TILE_BYTES;
  static constexpr int OFF_STG = OFF_DS + DS_BYTES;         // [2 groups]
  static constexpr int OFF_VEC = OFF_STG + 2 * STG_BYTES;
  static constexpr int OFF_BAR = OFF_VEC + VEC_BYTES;
  static constexpr int SMEM_USED = OFF_BAR + 256;
  static constexpr int SMEM_BYTES = (SMEM_USED + 1024 <= 232448) ? SMEM_USED + 1024 : 232448;
  static constexpr int T_S = 0, T_DP = 128, T_DV = 256, T_DK = 256 + D;
  static constexpr int NTHREADS = 384;
};

__device__ __forceinline__ void bulk_load_1d(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
      ::"r"(smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(gsrc)), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}

// Pre-pass of the tensor-core backward: D_i = sum_x dO*O and LSE_i (log2 domain), rows padded to
// Npad per (b,h) with +inf / 0 so that out-of-range queries contribute P = 0; the same kernel
// zero-fills the fp32 dQ accumulator (saves a separate memset pass).  16 lanes x 16 bytes cover one
// 128-element row, so every load is a full 128-bit coalesced access.
template <int D>
__global__ void __launch_bounds__(256)
    bwd_prep_tc_kernel(int B, int H, int N, int Npad, long long sb, long long sh, long long sn,
                       const __nv_bfloat16* __restrict__ O, const __nv_bfloat16* __restrict__ dO,
                       const float* __restrict__ M, const float* __restrict__ L, float* __restrict__ lse2,
                       float* __restrict__ dvec, float4* __restrict__ dq_acc4, long long n_acc4) {
  constexpr int LPR = D / 8;  // lanes per row (8 bf16 = 16 bytes each)
  const long long rows = static_cast<long long>(B) * H * Npad;
  const long long tid = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long nthreads = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long i = tid; i < n_acc4; i += nthreads) dq_acc4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  const int sub = threadIdx.x % LPR;
  for (long long r = tid / LPR; r < rows; r += nthreads / LPR) {
    const int n = static_cast<int>(r % Npad);
    const long long bh = r / Npad;
    float s = 0.f;
    if (n < N) {
      const long long off = (bh / H) * sb + (bh % H) * sh + static_cast<long long>(n) * sn + sub * 8;
      const uint4 a = __ldg(reinterpret_cast<const uint4*>(O + off));
      const uint4 c = __ldg(reinterpret_cast<const uint4*>(dO + off));
      const uint32_t aw[4] = {a.x, a.y, a.z, a.w}, cw[4] = {c.x, c.y, c.z, c.w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        s = fmaf(__uint_as_float(aw[k] << 16), __uint_as_float(cw[k] << 16), s);
        s = fmaf(__uint_as_float(aw[k] & 0xffff0000u), __uint_as_float(cw[k] & 0xffff0000u), s);
      }
    }
#pragma unroll
    for (int o = LPR / 2; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (sub == 0) {
      if (n < N) {
        const long long sr = bh * N + n;
        const float l = L[sr];
        lse2[r] = (l > 0.f) ? (M[sr] + logf(l)) * 1.4426950408889634f : INFINITY;
        dvec[r] = s;
      } else {
        lse2[r] = INFINITY;
        dvec[r] = 0.f;
      }
    }
  }
}

// dQ (bf16) = scale * dq_acc (fp32), contiguous accumulator -> strided output
__global__ void bwd_convert_dq_kernel(int H, int N, int D, long long sb, long long sh, long long sn, float scale,
                                      const float* __restrict__ acc, __nv_bfloat16* __restrict__ dQ, long long total8) {
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total8;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long e = i * 8;
    const int x = static_cast<int>(e % D);
    const long long row = e / D;
    const int n = static_cast<int>(row % N);
    const long long bh = row / N;
    const float4 a = __ldg(reinterpret_cast<const float4*>(acc + e));
    const float4 b = __ldg(reinterpret_cast<const float4*>(acc + e + 4));
    uint4 o;
    o.x = pack_bf16x2(a.x * scale, a.y * scale);
    o.y = pack_bf16x2(a.z * scale, a.w * scale);
    o.z = pack_bf16x2(b.x * scale, b.y * scale);
    o.w = pack_bf16x2(b.z * scale, b.w * scale);
    *reinterpret_cast<uint4*>(dQ + (bh / H) * sb + (bh % H) * sh + static_cast<long long>(n) * sn + x) = o;
  }
}

// exp2 of one 4-column group -> two packed bf16x2 registers
#define FA_BWD_P4(c4, MASKED)                                                              \
  {                                                                                        \
    const float4 l4 = *reinterpret_cast<const float4*>(lse + 4 * (c4));                    \
    float e0 = ex2_approx(fmaf(s[4 * (c4) + 0], p.scale_log2, -l4.x));                     \
    float e1 = ex2_approx(fmaf(s[4 * (c4) + 1], p.scale_log2, -l4.y));                     \
    float e2 = ex2_approx(fmaf(s[4 * (c4) + 2], p.scale_log2, -l4.z));                     \
    float e3 = ex2_approx(fmaf(s[4 * (c4) + 3], p.scale_log2, -l4.w));                     \
    if (MASKED) {                                                                          \
      if (!key_ok || 4 * (c4) + 0 < cmin) e0 = 0.f;                                        \
      if (!key_ok || 4 * (c4) + 1 < cmin) e1 = 0.f;                                        \
      if (!key_ok || 4 * (c4) + 2 < cmin) e2 = 0.f;                                        \
      if (!key_ok || 4 * (c4) + 3 < cmin) e3 = 0.f;                                        \
    }                                                                                      \
    pk[2 * (c4)] = pack_bf16x2(e0, e1);                                                    \
    pk[2 * (c4) + 1] = pack_bf16x2(e2, e3);                                                \
  }

__device__ __forceinline__ uint32_t bf16x2_mul(uint32_t a, uint32_t b) {
  uint32_t r;
  asm("mul.rn.bf16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
  return r;
}

template <int D, bool CAUSAL>
__global__ void __launch_bounds__(384, 1)
    bwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
               const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmdO,
               const __grid_constant__ CUtensorMap tmdQ, const BwdParams p) {
  using Cfg = BwdCfg<D>;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // align by offset (not by pointer cast) so the compiler keeps the shared address space: LDS/STS, not LD/ST
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  if (threadIdx.x == 0 && (smem - smem_raw) + Cfg::SMEM_USED > Cfg::SMEM_BYTES) {
    printf("fa bwd: dynamic shared memory base misaligned by %d bytes\n", (int)(smem - smem_raw));
    __trap();
  }
  uint8_t* sK = smem + Cfg::OFF_K;
  uint8_t* sV = smem + Cfg::OFF_V;
  uint8_t* sQ = smem + Cfg::OFF_Q;
  uint8_t* sdO = smem + Cfg::OFF_DO;
  uint8_t* sdS = smem + Cfg::OFF_DS;
  uint8_t* sStg = smem + Cfg::OFF_STG;
  float* sVec = reinterpret_cast<float*>(smem + Cfg::OFF_VEC);  // [stage][2][128]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::OFF_BAR);
  uint64_t* kv_full = bars;          // [1]
  uint64_t* q_full = bars + 1;       // [2]  Q_i + LSE_i + D_i
  uint64_t* q_empty = bars + 3;      // [2]
  uint64_t* do_full = bars + 5;      // [1]  dO_i (single stage)
  uint64_t* do_empty = bars + 6;     // [1]
  uint64_t* s_full = bars + 7;       // [2]  (index = iteration parity)
  uint64_t* p_full = bars + 9;       // [2]
  uint64_t* dp_full = bars + 11;     // [2]
  uint64_t* ds_full = bars + 13;     // [2]
  uint64_t* ds_empty = bars + 15;    // [2]
  uint64_t* dq_full = bars + 17;     // [2]
  uint64_t* dq_free = bars + 19;     // [2]
  uint64_t* dkv_done = bars + 21;    // [1]
  uint64_t* stg_full = bars + 22;    // [2 buffers]        staging box written by a compute group
  uint64_t* stg_empty = bars + 24;   // [2 groups][2 buf]  staging box read by the TMA reduction
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 28);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int kt = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int k0 = kt * 128;
#ifdef FA_TRACE
  long long* tr = (blockIdx.x == 1 && blockIdx.y == 0 && blockIdx.z == 0 && lane == 0) ? p.trace : nullptr;
#endif
  int kv_end = p.N;
  if (p.kv_len) kv_end = min(kv_end, max(__ldg(p.kv_len + b), 0));
  const int nq = (p.N + 127) >> 7;
  const int q_first = CAUSAL ? kt : 0;
  const int n_iter = nq - q_first;
  __nv_bfloat16* dKb = reinterpret_cast<__nv_bfloat16*>(p.dK) + b * p.sb + h * p.sh;
  __nv_bfloat16* dVb = reinterpret_cast<__nv_bfloat16*>(p.dV) + b * p.sb + h * p.sh;

  if (k0 >= kv_end) {
    // every key of this tile is padding: its gradients are exactly zero
    for (int idx = threadIdx.x; idx < 128 * (D / 8); idx += blockDim.x) {
      const int r = idx / (D / 8), c = (idx % (D / 8)) * 8;
      if (k0 + r < p.N) {
        const uint4 z = make_uint4(0, 0, 0, 0);
        *reinterpret_cast<uint4*>(dKb + static_cast<long long>(k0 + r) * p.sn + c) = z;
        *reinterpret_cast<uint4*>(dVb + static_cast<long long>(k0 + r) * p.sn + c) = z;
      }
    }
    return;
  }

  if (warp == 8 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    tma_prefetch_desc(&tmdO);
    tma_prefetch_desc(&tmdQ);
    mbar_init(kv_full, 1);
    mbar_init(dkv_done, 1);
    mbar_init(do_full, 1);
    mbar_init(do_empty, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&q_full[i], 1);
      mbar_init(&q_empty[i], 1);
      mbar_init(&s_full[i], 1);
      mbar_init(&p_full[i], 128);
      mbar_init(&dp_full[i], 1);
      mbar_init(&ds_full[i], 128);
      mbar_init(&ds_empty[i], 1);
      mbar_init(&dq_full[i], 1);
      mbar_init(&dq_free[i], 128);
      mbar_init(&stg_full[i], 128);
      mbar_init(&stg_empty[i], 1);
      mbar_init(&stg_empty[2 + i], 1);
    }
    fence_mbar_init();
  }
  if (warp == 9) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp >= 8) {
    reg_dealloc<88>();
    if (warp == 8) {
      // ---------------------------------------------------------------- TMA producer
      if (lane == 0) {
        mbar_expect_tx(kv_full, 2 * Cfg::TILE_BYTES);
#pragma unroll
        for (int c = 0; c < Cfg::NCHUNK; ++c) {
          tma_load_4d(sK + c * Cfg::CHUNK_BYTES, &tmK, kv_full, c * 64, k0, h, b);
          tma_load_4d(sV + c * Cfg::CHUNK_BYTES, &tmV, kv_full, c * 64, k0, h, b);
        }
        const long long vec_base = (static_cast<long long>(b) * p.H + h) * p.Npad;
        for (int it = 0; it < n_iter; ++it) {
          const int s = it & 1;
          const int q0 = (q_first + it) * 128;
          mbar_wait(&q_empty[s], ((it >> 1) & 1) ^ 1);
          mbar_expect_tx(&q_full[s], Cfg::TILE_BYTES + 1024);
#pragma unroll
          for (int c = 0; c < Cfg::NCHUNK; ++c)
            tma_load_4d(sQ + s * Cfg::TILE_BYTES + c * Cfg::CHUNK_BYTES, &tmQ, &q_full[s], c * 64, q0, h, b);
          bulk_load_1d(sVec + s * 256, p.lse2 + vec_base + q0, 512, &q_full[s]);
          bulk_load_1d(sVec + s * 256 + 128, p.dvec + vec_base + q0, 512, &q_full[s]);
          mbar_wait(do_empty, (it & 1) ^ 1);
          mbar_expect_tx(do_full, Cfg::TILE_BYTES);
#pragma unroll
          for (int c = 0; c < Cfg::NCHUNK; ++c)
            tma_load_4d(sdO + c * Cfg::CHUNK_BYTES, &tmdO, do_full, c * 64, q0, h, b);
        }
      }
      __syncwarp();
    } else if (warp == 9) {
      // ---------------------------------------------------------------- MMA issuer
      // The whole warp walks the schedule (uniform control flow keeps the descriptors in uniform
      // registers: no per-lane serialisation loop around every tcgen05.mma); one elected lane issues.
      {
        const bool leader = elect_one();
        constexpr uint32_t idesc_s = make_idesc_bf16(128, 128, 0, 0);
        constexpr uint32_t idesc_kn = make_idesc_bf16(128, D, 0, 1);   // A K-major / TMEM, B MN-major
        constexpr uint32_t idesc_mn = make_idesc_bf16(128, D, 1, 1);   // A MN-major, B MN-major
        const uint32_t tS = tmem_base + Cfg::T_S, tdP = tmem_base + Cfg::T_DP;
        const uint32_t tdV = tmem_base + Cfg::T_DV, tdK = tmem_base + Cfg::T_DK;
        // descriptor low words (address >> 4 | LBO) per tile and the two high words (K-major / MN-major)
        const uint64_t dkm = make_smem_desc(smem_u32(sK), 16, 1024);
        const uint64_t dmn = make_smem_desc(smem_u32(sK), Cfg::CHUNK_BYTES, 1024);
        const uint32_t km_hi = static_cast<uint32_t>(dkm >> 32), mn_hi = static_cast<uint32_t>(dmn >> 32);
        const uint32_t km_lo0 = static_cast<uint32_t>(dkm), mn_lo0 = static_cast<uint32_t>(dmn);
        auto km_lo = [&](const uint8_t* ptr) { return km_lo0 + ((smem_u32(ptr) - smem_u32(sK)) >> 4); };
        auto mn_lo = [&](const uint8_t* ptr) { return mn_lo0 + ((smem_u32(ptr) - smem_u32(sK)) >> 4); };
        const uint32_t kK = km_lo(sK), kV = km_lo(sV), kdO = km_lo(sdO), kDS = km_lo(sdS);
        const uint32_t mK = mn_lo(sK), mdO = mn_lo(sdO), mDS = mn_lo(sdS);
        auto commit = [&](uint64_t* bar) {
          if (leader) mma_commit(bar);
        };
        // [128 x 128] = X[128 x D] * Y[128 x D]^T : both operands K-major over the head dim
        auto issue_nt = [&](uint32_t dst, uint32_t xa, uint32_t ya) {
          if (leader) {
#pragma unroll
            for (int k = 0; k < D / 16; ++k) {
              const uint32_t off = ((k >> 2) * Cfg::CHUNK_BYTES + (k & 3) * 32) >> 4;
              mma_ss2(dst, xa + off, km_hi, ya + off, km_hi, idesc_s, k > 0);
            }
          }
        };
        mbar_wait(kv_full, 0);
        mbar_wait(&q_full[0], 0);
        tc_fence_after();
        issue_nt(tS, kK, km_lo(sQ));
        commit(&s_full[0]);
        mbar_wait(do_full, 0);
        tc_fence_after();
        issue_nt(tdP, kV, kdO);
        commit(&dp_full[0]);
        for (int it = 0; it < n_iter; ++it) {
          const int s = it & 1, ph = (it >> 1) & 1;
          const uint32_t mQ = mn_lo(sQ + s * Cfg::TILE_BYTES);
          // dV += P^T dO_i : A = P^T in TMEM (bf16, 8 columns per 16 queries), B = dO_i as [K=q][N=d]
          mbar_wait(&p_full[s], ph);
          tc_fence_after();
          FA_TR(0)
          if (leader) {
#pragma unroll
            for (int k = 0; k < 8; ++k)
              mma_ts2(tdV, tS + k * 8, mdO + k * (2048 >> 4), mn_hi, idesc_kn, (it > 0 || k > 0) ? 1u : 0u);
          }
          commit(do_empty);   // dO_i is dead once dV(i) has run (dP(i) ran earlier)
          FA_TR(1)
          if (it + 1 < n_iter) {
            const int s1 = (it + 1) & 1;
            mbar_wait(&q_full[s1], ((it + 1) >> 1) & 1);
            tc_fence_after();
            FA_TR(2)
            issue_nt(tS, kK, km_lo(sQ + s1 * Cfg::TILE_BYTES));
            commit(&s_full[s1]);
            FA_TR(3)
          }
          mbar_wait(&ds_full[s], ph);
          tc_fence_after();
          FA_TR(4)
          if (leader) {
            // dQ_i = dS K : A = dS^T tile read MN-major ([K=key][M=q]), B = K tile as [K=key][N=d]
#pragma unroll
            for (int k = 0; k < 8; ++k)
              mma_ss2(tdP, mDS + k * (2048 >> 4), mn_hi, mK + k * (2048 >> 4), mn_hi, idesc_mn, k > 0);
          }
          commit(&dq_full[s]);
          if (leader) {
            // dK += dS^T Q_i : A = dS^T tile K-major ([M=key][K=q]), B = Q_i as [K=q][N=d]
#pragma unroll
            for (int k = 0; k < 8; ++k)
              mma_ss2(tdK, kDS + (((k >> 2) * Cfg::CHUNK_BYTES + (k & 3) * 32) >> 4), km_hi, mQ + k * (2048 >> 4),
                      mn_hi, idesc_kn, (it > 0 || k > 0) ? 1u : 0u);
          }
          commit(&ds_empty[s]);
          commit(&q_empty[s]);
          FA_TR(5)
          if (it + 1 < n_iter) {
            const int s1 = (it + 1) & 1;
            mbar_wait(do_full, (it + 1) & 1);
            mbar_wait(&dq_free[s], ph);
            tc_fence_after();
            FA_TR(6)
            issue_nt(tdP, kV, kdO);
            commit(&dp_full[s1]);
            FA_TR(7)
          }
        }
        commit(dkv_done);
      }
      __syncwarp();
    } else if (warp == 10) {
      // ---------------------------------------------------------------- dQ reduction issuer
      // Rounds (it, c): the compute group of iteration `it` fills staging buffer c&1 with columns
      // [32c, 32c+32) of dQ_it; this thread turns each into one TMA add-reduction and hands the
      // buffer back (to whichever group uses it two rounds later) once the TMA has read it.
      if (lane == 0) {
        constexpr int NR = D / 32;
        int prev_g = 0, prev_b = 0;
        bool have_prev = false;
        for (int it = 0; it < n_iter; ++it) {
          const int q0 = (q_first + it) * 128;
          for (int c = 0; c < NR; ++c) {
            const int bsel = c & 1;
            mbar_wait(&stg_full[bsel], ((it * (NR / 2) + (c >> 1)) & 1));
            tma_reduce_add_4d(&tmdQ, sStg + bsel * Cfg::STG_BYTES, 32 * c, q0, h, b);
            tma_store_commit();
            if (have_prev) {
              tma_store_wait_read<1>();                       // the previous round's box has been read
              mbar_arrive(&stg_empty[2 * prev_g + prev_b]);
            }
            // next user of this buffer: same iteration if c+2 < NR, otherwise the other group
            prev_g = (c + 2 < NR) ? (it & 1) : ((it + 1) & 1);
            prev_b = bsel;
            have_prev = true;
          }
        }
        tma_store_wait_all<0>();
      }
      __syncwarp();
    }
  } else {
    // ------------------------------------------------------------------ compute groups
    reg_alloc<208>();
    const int g = warp >> 2, w = warp & 3;
    const int j = w * 32 + lane;                 // this thread's TMEM lane: key row (S^T, dP^T) / query row (dQ)
    const uint32_t lane_base = tmem_base + (static_cast<uint32_t>(w * 32) << 16);
    const uint32_t tS = lane_base + Cfg::T_S, tdP = lane_base + Cfg::T_DP;
    const bool key_ok = (k0 + j) < kv_end;
    const bool all_keys_ok = (k0 + 128) <= kv_end;
    uint8_t* ds_row = sdS + j * 128;
    uint8_t* stg_row = sStg + j * 128;
    const int jx = j & 7;
#ifdef FA_TRACE
    long long* trw = tr;   // lane 0 of every warp of the group
    if (w != 0) tr = nullptr;
#endif

    for (int it = g; it < n_iter; it += 2) {
      const int ph = (it >> 1) & 1;
      const int q0 = (q_first + it) * 128;
      const float* lse = sVec + g * 256;
      const float* dv = lse + 128;
      // ---- P^T = exp2(S^T * scale*log2e - LSE2[q])
      mbar_wait(&s_full[g], ph);
      tc_fence_after();
      FA_TR(8)
      FA_TRW(2)
      float s[128];
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        uint32_t u[32];
        tmem_ld32(tS + 32 * c, u);
#pragma unroll
        for (int i = 0; i < 32; ++i) s[32 * c + i] = __uint_as_float(u[i]);
      }
      tmem_wait_ld();
      FA_TR(9)
      mbar_wait(&q_full[g], ph);  // LSE / D vectors of this Q tile are in shared memory
      uint32_t pk[64];
      const bool diag = CAUSAL && (q0 < k0 + 128);
      const int cmin = CAUSAL ? (k0 + j - q0) : 0;  // causal: queries q0+c < key are masked
      if (all_keys_ok && !diag) {
#pragma unroll
        for (int c4 = 0; c4 < 32; ++c4) FA_BWD_P4(c4, false)
      } else {
#pragma unroll
        for (int c4 = 0; c4 < 32; ++c4) FA_BWD_P4(c4, true)
      }
#pragma unroll
      for (int c = 0; c < 2; ++c) tmem_st32(tS + 32 * c, *reinterpret_cast<uint32_t(*)[32]>(&pk[32 * c]));
      tmem_wait_st();
      tc_fence_before();
      mbar_arrive(&p_full[g]);
      FA_TR(10)
      FA_TRW(0)

      // ---- dS^T = P^T * (dP^T - D[q])  -> shared memory (bf16, 128B-swizzled rows)
      mbar_wait(&dp_full[g], ph);
      tc_fence_after();
      FA_TR(11)
      FA_TRW(3)
      if (it >= 1) mbar_wait(&ds_empty[g ^ 1], ((it - 1) >> 1) & 1);  // dK(it-1) done with the buffer
      uint32_t ua[2][32];
      tmem_ld32(tdP, ua[0]);
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        tmem_wait_ld();                                         // chunk c is in registers
        if (c < 3) tmem_ld32(tdP + 32 * (c + 1), ua[(c + 1) & 1]);  // prefetch the next 32 columns
        if (c == 3) tc_fence_before();
        const uint32_t(&u)[32] = ua[c & 1];
#pragma unroll
        for (int v8 = 0; v8 < 4; ++v8) {   // 8 queries -> one 16-byte piece
          const float4 da = *reinterpret_cast<const float4*>(dv + 32 * c + 8 * v8);
          const float4 db = *reinterpret_cast<const float4*>(dv + 32 * c + 8 * v8 + 4);
          uint4 o;
          o.x = bf16x2_mul(pk[16 * c + 4 * v8 + 0], pack_bf16x2(__uint_as_float(u[8 * v8 + 0]) - da.x,
                                                                __uint_as_float(u[8 * v8 + 1]) - da.y));
          o.y = bf16x2_mul(pk[16 * c + 4 * v8 + 1], pack_bf16x2(__uint_as_float(u[8 * v8 + 2]) - da.z,
                                                                __uint_as_float(u[8 * v8 + 3]) - da.w));
          o.z = bf16x2_mul(pk[16 * c + 4 * v8 + 2], pack_bf16x2(__uint_as_float(u[8 * v8 + 4]) - db.x,
                                                                __uint_as_float(u[8 * v8 + 5]) - db.y));
          o.w = bf16x2_mul(pk[16 * c + 4 * v8 + 3], pack_bf16x2(__uint_as_float(u[8 * v8 + 6]) - db.z,
                                                                __uint_as_float(u[8 * v8 + 7]) - db.w));
          const int unit = 4 * (c & 1) + v8;  // 16-byte unit inside the 128-byte row of this chunk
          *reinterpret_cast<uint4*>(ds_row + (c >> 1) * Cfg::CHUNK_BYTES + ((unit ^ jx) << 4)) = o;
        }
      }
      fence_proxy_async_smem();
      mbar_arrive(&ds_full[g]);
      FA_TR(12)
      FA_TRW(1)

      // ---- drain dQ_i (lane = query row): TMEM -> registers -> swizzled staging box -> TMA add-reduce
      mbar_wait(&dq_full[g], ph);
      tc_fence_after();
      FA_TR(13)
      uint32_t dq[D / 32][32];
#pragma unroll
      for (int c = 0; c < D / 32; ++c) tmem_ld32(tdP + 32 * c, dq[c]);
      tmem_wait_ld();
      tc_fence_before();
      mbar_arrive(&dq_free[g]);          // T_dP may be overwritten by dP(it+1)
      FA_TR(14)
#pragma unroll
      for (int c = 0; c < D / 32; ++c) {
        // use number u of buffer c&1 by THIS group; group 0's very first use finds the buffer free
        const int u = (it >> 1) * (D / 64) + (c >> 1);
        mbar_wait(&stg_empty[2 * g + (c & 1)], g == 0 ? ((u & 1) ^ 1) : (u & 1));
        uint8_t* dst = stg_row + (c & 1) * Cfg::STG_BYTES;
#pragma unroll
        for (int u8 = 0; u8 < 8; ++u8)
          *reinterpret_cast<uint4*>(dst + ((u8 ^ jx) << 4)) =
              make_uint4(dq[c][4 * u8], dq[c][4 * u8 + 1], dq[c][4 * u8 + 2], dq[c][4 * u8 + 3]);
        fence_proxy_async_smem();
        mbar_arrive(&stg_full[c & 1]);
      }
      FA_TR(15)
    }

    // ---- epilogue: group 0 stores dK (scaled), group 1 stores dV
    mbar_wait(dkv_done, 0);
    tc_fence_after();
    const uint32_t tacc = lane_base + (g == 0 ? Cfg::T_DK : Cfg::T_DV);
    const float mul = (g == 0) ? p.scale : 1.0f;
    __nv_bfloat16* orow = (g == 0 ? dKb : dVb) + static_cast<long long>(k0 + j) * p.sn;
#pragma unroll
    for (int c = 0; c < D / 32; ++c) {
      uint32_t u[32];
      tmem_ld32(tacc + 32 * c, u);
      tmem_wait_ld();
      if (k0 + j < p.N) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          uint4 o;
          o.x = pack_bf16x2(__uint_as_float(u[8 * i]) * mul, __uint_as_float(u[8 * i + 1]) * mul);
          o.y = pack_bf16x2(__uint_as_float(u[8 * i + 2]) * mul, __uint_as_float(u[8 * i + 3]) * mul);
          o.z = pack_bf16x2(__uint_as_float(u[8 * i + 4]) * mul, __uint_as_float(u[8 * i + 5]) * mul);
          o.w = pack_bf16x2(__uint_as_float(u[8 * i + 6]) * mul, __uint_as_float(u[8 * i + 7]) * mul);
          *reinterpret_cast<uint4*>(orow + 32 * c + 8 * i) = o;
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 9) tmem_dealloc<512>(tmem_base);
}

}  // namespace sm100
}  // namespace fa
