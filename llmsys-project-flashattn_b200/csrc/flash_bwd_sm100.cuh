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
//   warps 0-15      : four compute warpgroups (g, hh).  The two groups g alternate Q tiles (ping-pong);
//                     inside a group two threads share a key row (TMEM lane), hh picks 64 of the 128
//                     query columns: exp2 -> P^T, dS^T = P^T*(dP^T - D), then drain half of dQ_i:
//                     TMEM -> registers -> 128B-swizzled staging box in shared memory ->
//                     cp.reduce.async.bulk.tensor (fp32 add) into the dQ accumulator (LSU red.global
//                     tops out near 9 clk per warp instruction on B200, far too slow for 64 KB per
//                     tile pair).  T_dP is shared by dP(i), dS(i) and dQ(i), so the chain
//                     dP(i) -> dS(i) -> dQ(i) -> drain -> dP(i+1) is the critical cycle of the kernel;
//                     two threads per row halve every compute link of it.  No exchange is needed
//                     between the halves: the backward uses the saved LSE, and each thread keeps its
//                     P^T inside its own S^T columns (the dV GEMM reads the two 32-column pieces).
//   warp 16         : TMA producer (K,V once; Q_i + LSE_i + D_i 2-stage ring; dO_i single stage)
//   warp 17         : tcgen05.mma issuer + TMEM allocation
//   warps 18, 19    : issue the dQ TMA reductions of the hh = 0 / hh = 1 warpgroups, one staging box each
//   (20 warps launch with 96 registers; setmaxnreg moves the compute warpgroups to 104 and the helper
//    warpgroup to 64)
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
  int d;               // real head dim (<= D): columns d..D-1 are TMA zero-fill (inputs) / never stored (outputs)
  const int* kv_len;
  const float* key_mask;   // (B, N) additive mask or nullptr (template MASK2)
  const float* lse2;   // (B*H, Npad) NEGATED log2-domain LSE (an FFMA2 addend); -inf for rows >= N
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

#ifndef FA_BWD_DK_TS
#define FA_BWD_DK_TS 0   // 1: dS^T (bf16) is also written over the thread's own dP^T columns in TMEM and the dK GEMM takes
                         // its A operand from there (TS) instead of shared memory: -256 MMA-operand wavefronts per tile pair,
                         // at the price of issuing dK before dQ (dQ overwrites those columns)
#endif
#ifndef FA_BWD_DQ_RED
#define FA_BWD_DQ_RED 0   // 0: dQ through shared-memory staging + TMA add-reduce; 1 / 2: red.global.v4 from registers
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
  static constexpr int NTHREADS = 640;
};

__device__ __forceinline__ void bulk_load_1d(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
      ::"r"(smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(gsrc)), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}

// Pre-pass of the tensor-core backward: D_i = sum_x dO*O and LSE_i (log2 domain), rows padded to
// Npad per (b,h) with -inf (the LSE is stored negated) / 0 so that out-of-range queries contribute P = 0; the same kernel
// zero-fills the fp32 dQ accumulator (saves a separate memset pass).  16 lanes x 16 bytes cover one
// 128-element row, so every load is a full 128-bit coalesced access.
template <int D>
__global__ void __launch_bounds__(256)
    bwd_prep_tc_kernel(int B, int H, int N, int Npad, int d_real, long long sb, long long sh, long long sn,
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
    if (n < N && sub * 8 < d_real) {
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
        lse2[r] = (l > 0.f) ? -(M[sr] + logf(l)) * 1.4426950408889634f : -INFINITY;
        dvec[r] = s;
      } else {
        lse2[r] = -INFINITY;
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

__device__ __forceinline__ uint32_t bf16x2_mul(uint32_t a, uint32_t b) {
  uint32_t r;
  asm("mul.rn.bf16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
  return r;
}

template <int D, bool CAUSAL, bool MASK2>
__global__ void __launch_bounds__(640, 1)
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
  uint64_t* stg_full = bars + 22;    // [2 halves]            staging box hh written by warpgroup (g, hh)
  uint64_t* stg_empty = bars + 24;   // [2 halves][2 groups]  staging box hh read by the TMA reduction, free for (g, hh)
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
      if (k0 + r < p.N && c < p.d) {
        const uint4 z = make_uint4(0, 0, 0, 0);
        *reinterpret_cast<uint4*>(dKb + static_cast<long long>(k0 + r) * p.sn + c) = z;
        *reinterpret_cast<uint4*>(dVb + static_cast<long long>(k0 + r) * p.sn + c) = z;
      }
    }
    return;
  }

  if (warp == 16 && lane == 0) {
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
      mbar_init(&p_full[i], 256);
      mbar_init(&dp_full[i], 1);
      mbar_init(&ds_full[i], 256);
      mbar_init(&ds_empty[i], 1);
      mbar_init(&dq_full[i], 1);
      mbar_init(&dq_free[i], 256);
      mbar_init(&stg_full[i], 128);
      mbar_init(&stg_empty[i], 1);
      mbar_init(&stg_empty[2 + i], 1);
    }
    fence_mbar_init();
  }
  if (warp == 17) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp >= 16) {
    reg_dealloc<64>();
    if (warp == 16) {
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
    } else if (warp == 17) {
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
          // dV += P^T dO_i : A = P^T in TMEM (bf16, 8 columns per 16 queries; queries [0,64) sit in columns
          // [0,32) and queries [64,128) in columns [64,96) of T_S), B = dO_i as [K=q][N=d]
          mbar_wait(&p_full[s], ph);
          tc_fence_after();
          FA_TR(0)
          if (leader) {
#pragma unroll
            for (int k = 0; k < 8; ++k)
              mma_ts2(tdV, tS + (k >> 2) * 64 + (k & 3) * 8, mdO + k * (2048 >> 4), mn_hi, idesc_kn,
                      (it > 0 || k > 0) ? 1u : 0u);
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
#if FA_BWD_DK_TS
          if (leader) {
            // dK += dS^T Q_i : A = dS^T straight from TMEM (bf16 over the dP^T columns, same packing as P^T), B = Q_i
#pragma unroll
            for (int k = 0; k < 8; ++k)
              mma_ts2(tdK, tdP + (k >> 2) * 64 + (k & 3) * 8, mQ + k * (2048 >> 4), mn_hi, idesc_kn,
                      (it > 0 || k > 0) ? 1u : 0u);
            // dQ_i = dS K : overwrites T_dP (the tensor pipe runs in issue order: dK has read dS^T by then)
#pragma unroll
            for (int k = 0; k < 8; ++k)
              mma_ss2(tdP, mDS + k * (2048 >> 4), mn_hi, mK + k * (2048 >> 4), mn_hi, idesc_mn, k > 0);
          }
          commit(&dq_full[s]);
#else
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
#endif
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
    } else {
      // ---------------------------------------------------------------- dQ reduction issuers (warps 18, 19)
      // Warp 18 + hh serves the warpgroups (., hh): staging box hh receives columns [D/2*hh + 32*c2, +32) of
      // dQ_it from warpgroup (it & 1, hh); one lane turns each fill into a TMA add-reduction and hands the
      // box to its next user once the TMA has read it.  The two halves never wait for each other.
#if !defined(FA_EXP_NODRAIN) && !FA_BWD_DQ_RED
      if (lane == 0) {
        constexpr int PER_WG = D / 64;   // 32-column boxes per warpgroup and iteration
        const int hh = warp - 18;
        for (int it = 0; it < n_iter; ++it) {
          const int q0 = (q_first + it) * 128;
          for (int c2 = 0; c2 < PER_WG; ++c2) {
            mbar_wait(&stg_full[hh], (it * PER_WG + c2) & 1);
            tma_reduce_add_4d(&tmdQ, sStg + hh * Cfg::STG_BYTES, (D / 2) * hh + 32 * c2, q0, h, b);
            tma_store_commit();
            tma_store_wait_read<0>();                         // the box has been read
            // next user: the same warpgroup for its next box of this iteration, otherwise the other group
            mbar_arrive(&stg_empty[2 * hh + ((c2 + 1 < PER_WG) ? (it & 1) : ((it + 1) & 1))]);
          }
        }
        tma_store_wait_all<0>();
      }
#endif
      __syncwarp();
    }
  } else {
    // ------------------------------------------------------------------ compute warpgroups
    reg_alloc<104>();
    const int wg = warp >> 2, w = warp & 3;
    const int g = wg & 1, hh = wg >> 1;          // Q-tile parity group, query-column half
    const int j = w * 32 + lane;                 // this thread's TMEM lane: key row (S^T, dP^T) / query row (dQ)
    const uint32_t lane_base = tmem_base + (static_cast<uint32_t>(w * 32) << 16);
    const uint32_t tS = lane_base + Cfg::T_S + 64 * hh;         // my 64 S^T columns; my P^T = the first 32 of them
    const uint32_t tdP = lane_base + Cfg::T_DP + 64 * hh;       // my 64 dP^T columns
    const uint32_t tdQ = lane_base + Cfg::T_DP + (D / 2) * hh;  // my D/2 dQ columns
    const bool key_ok = (k0 + j) < kv_end;
    const bool all_keys_ok = (k0 + 128) <= kv_end;
    uint8_t* ds_row = sdS + hh * Cfg::CHUNK_BYTES + j * 128;    // my 64 queries = one 128-byte swizzled row
    uint8_t* stg_row = sStg + j * 128;
    const int jx = j & 7;
    const uint64_t sc2 = f32x2(p.scale_log2, p.scale_log2);
    // generic additive key mask: in the transposed layout a thread owns ONE key, so the mask is a per-thread scalar.
    // It is added to the scaled score BEFORE the (negated) LSE, the order the forward uses (a -1e8 "padding" mask
    // swallows the score in fp32 there, too -- the two passes have to round alike).
    float mk = 0.f;
    if (MASK2 && (k0 + j) < p.N) mk = __ldg(p.key_mask + static_cast<long long>(b) * p.N + k0 + j) * 1.4426950408889634f;
    const uint64_t mk2 = f32x2(mk, mk);
#ifdef FA_TRACE
    long long* trw = (hh == 0) ? tr : nullptr;   // lane 0 of every warp of the hh = 0 warpgroups
    if (w != 0 || hh != 0) tr = nullptr;
#endif

    for (int it = g; it < n_iter; it += 2) {
      const int ph = (it >> 1) & 1;
      const int q0 = (q_first + it) * 128;
      const float* nlse = sVec + g * 256 + 64 * hh;   // -LSE2 of my 64 queries
      const float* dv = sVec + g * 256 + 128 + 64 * hh;
      // ---- P^T = exp2(S^T * scale*log2e - LSE2[q])
      mbar_wait(&s_full[g], ph);
      tc_fence_after();
      FA_TR(8)
      FA_TRW(2)
      uint32_t pk[32];
      {
        float s[64];
        tmem_ld32f(tS, &s[0]);
        tmem_ld32f(tS + 32, &s[32]);
        tmem_wait_ld();
        FA_TR(9)
        mbar_wait(&q_full[g], ph);  // LSE / D vectors of this Q tile are in shared memory
        const bool diag = CAUSAL && (q0 < k0 + 128);
        const int cmin = CAUSAL ? (k0 + j - q0 - 64 * hh) : 0;  // causal: my queries with index < cmin precede the key
        const bool masked = !(all_keys_ok && !diag);
#pragma unroll
        for (int c4 = 0; c4 < 16; ++c4) {
#ifdef FA_EXP_NOSTATS   // timing-only experiment: no shared-memory loads of the row statistics
          const float4 l4 = make_float4(p.scale, p.scale, p.scale, p.scale);
#else
          const float4 l4 = *reinterpret_cast<const float4*>(nlse + 4 * c4);
#endif
          float x0, x1, x2, x3;
          if (MASK2) {
            f32x2_unpack(add_f32x2(fma_f32x2(f32x2(s[4 * c4], s[4 * c4 + 1]), sc2, mk2), f32x2(l4.x, l4.y)), x0, x1);
            f32x2_unpack(add_f32x2(fma_f32x2(f32x2(s[4 * c4 + 2], s[4 * c4 + 3]), sc2, mk2), f32x2(l4.z, l4.w)), x2, x3);
          } else {
            f32x2_unpack(fma_f32x2(f32x2(s[4 * c4], s[4 * c4 + 1]), sc2, f32x2(l4.x, l4.y)), x0, x1);
            f32x2_unpack(fma_f32x2(f32x2(s[4 * c4 + 2], s[4 * c4 + 3]), sc2, f32x2(l4.z, l4.w)), x2, x3);
          }
          float e0 = ex2_approx(x0), e1 = ex2_approx(x1), e2 = ex2_approx(x2), e3 = ex2_approx(x3);
          if (masked) {
            if (!key_ok || 4 * c4 + 0 < cmin) e0 = 0.f;
            if (!key_ok || 4 * c4 + 1 < cmin) e1 = 0.f;
            if (!key_ok || 4 * c4 + 2 < cmin) e2 = 0.f;
            if (!key_ok || 4 * c4 + 3 < cmin) e3 = 0.f;
          }
          pk[2 * c4] = pack_bf16x2(e0, e1);
          pk[2 * c4 + 1] = pack_bf16x2(e2, e3);
        }
      }
      tmem_st32(tS, pk);
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
      uint32_t ua[2][16];
      tmem_ld16(tdP, ua[0]);
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        tmem_wait_ld();                                          // 16-column piece c is in registers
        if (c < 3) tmem_ld16(tdP + 16 * (c + 1), ua[(c + 1) & 1]);   // prefetch the next piece
        const uint32_t(&u)[16] = ua[c & 1];
#pragma unroll
        for (int v8 = 0; v8 < 2; ++v8) {   // 8 queries -> one 16-byte piece
#ifdef FA_EXP_NOSTATS
          const float4 da = make_float4(p.scale, p.scale, p.scale, p.scale), db = da;
#else
          const float4 da = *reinterpret_cast<const float4*>(dv + 16 * c + 8 * v8);
          const float4 db = *reinterpret_cast<const float4*>(dv + 16 * c + 8 * v8 + 4);
#endif
          uint4 o;
          o.x = bf16x2_mul(pk[8 * c + 4 * v8 + 0], pack_bf16x2(__uint_as_float(u[8 * v8 + 0]) - da.x,
                                                               __uint_as_float(u[8 * v8 + 1]) - da.y));
          o.y = bf16x2_mul(pk[8 * c + 4 * v8 + 1], pack_bf16x2(__uint_as_float(u[8 * v8 + 2]) - da.z,
                                                               __uint_as_float(u[8 * v8 + 3]) - da.w));
          o.z = bf16x2_mul(pk[8 * c + 4 * v8 + 2], pack_bf16x2(__uint_as_float(u[8 * v8 + 4]) - db.x,
                                                               __uint_as_float(u[8 * v8 + 5]) - db.y));
          o.w = bf16x2_mul(pk[8 * c + 4 * v8 + 3], pack_bf16x2(__uint_as_float(u[8 * v8 + 6]) - db.z,
                                                               __uint_as_float(u[8 * v8 + 7]) - db.w));
          const int unit = 2 * c + v8;  // 16-byte unit inside my 128-byte row
          *reinterpret_cast<uint4*>(ds_row + ((unit ^ jx) << 4)) = o;
#if FA_BWD_DK_TS
          pk[8 * c + 4 * v8 + 0] = o.x, pk[8 * c + 4 * v8 + 1] = o.y;   // P^T of these queries is consumed: keep dS^T
          pk[8 * c + 4 * v8 + 2] = o.z, pk[8 * c + 4 * v8 + 3] = o.w;
#endif
        }
      }
#if FA_BWD_DK_TS
      tmem_st32(tdP, pk);          // dS^T (bf16) over my own dP^T columns [64hh, 64hh+32): A operand of the dK GEMM
      tmem_wait_st();
#endif
      tc_fence_before();
      fence_proxy_async_smem();
      mbar_arrive(&ds_full[g]);
      FA_TR(12)
      FA_TRW(1)

      // ---- drain my half of dQ_i (lane = query row): TMEM -> registers -> swizzled staging box -> TMA add-reduce
      mbar_wait(&dq_full[g], ph);
      tc_fence_after();
      FA_TR(13)
      uint32_t dq[D / 64][32];
#pragma unroll
      for (int c2 = 0; c2 < D / 64; ++c2) tmem_ld32(tdQ + 32 * c2, dq[c2]);
      tmem_wait_ld();
      tc_fence_before();
      mbar_arrive(&dq_free[g]);          // T_dP may be overwritten by dP(it+1)
      FA_TR(14)
#if FA_BWD_DQ_RED
      // dQ_i straight from registers into the fp32 accumulator with vector reductions (no shared-memory staging, no
      // TMA): a thread owns 64 consecutive floats of its query row.  With FA_BWD_DQ_RED == 2 lane pairs swap half of
      // their float4s first so that the two lanes of a pair complete one 32-byte sector per instruction.
      {
        const long long rowi = (static_cast<long long>(b) * p.H + h) * p.N + q0 + j;
        float* dst = p.dq_acc + rowi * D + (D / 2) * hh;
        const bool row_ok = (q0 + j) < p.N;
#if FA_BWD_DQ_RED == 2
        const bool odd = (lane & 1) != 0;
        const long long prow = odd ? -static_cast<long long>(D) : static_cast<long long>(D);   // partner's row
        const bool prow_ok = (q0 + (j ^ 1)) < p.N;
#pragma unroll
        for (int c2 = 0; c2 < D / 64; ++c2) {
#pragma unroll
          for (int s8 = 0; s8 < 4; ++s8) {   // a 32-byte sector = float4 pair (2*s8, 2*s8+1) of this 32-column piece
            uint32_t mine[4], give[4], got[4];
#pragma unroll
            for (int x = 0; x < 4; ++x) {
              // even lane keeps its LOW float4 and gives its HIGH one; odd lane keeps HIGH and gives LOW
              mine[x] = odd ? dq[c2][8 * s8 + 4 + x] : dq[c2][8 * s8 + x];
              give[x] = odd ? dq[c2][8 * s8 + x] : dq[c2][8 * s8 + 4 + x];
              got[x] = __shfl_xor_sync(0xffffffffu, give[x], 1);
            }
            // instruction 1: both lanes write the EVEN lane's row (even: own low half, odd: even's high half);
            // instruction 2: both lanes write the ODD lane's row (even: odd's low half, odd: own high half)
            float* a1 = dst + 32 * c2 + 8 * s8 + (odd ? (4 - D) : 0);      // even lane's row
            float* a2 = dst + 32 * c2 + 8 * s8 + (odd ? 4 : D);            // odd lane's row
            const bool ok1 = odd ? prow_ok : row_ok, ok2 = odd ? row_ok : prow_ok;
            const uint32_t* v1 = odd ? got : mine;
            const uint32_t* v2 = odd ? mine : got;
            (void)prow;
            if (ok1)
              asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(a1), "f"(__uint_as_float(v1[0])),
                           "f"(__uint_as_float(v1[1])), "f"(__uint_as_float(v1[2])), "f"(__uint_as_float(v1[3]))
                           : "memory");
            if (ok2)
              asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(a2), "f"(__uint_as_float(v2[0])),
                           "f"(__uint_as_float(v2[1])), "f"(__uint_as_float(v2[2])), "f"(__uint_as_float(v2[3]))
                           : "memory");
          }
        }
#else
        if (row_ok) {
#pragma unroll
          for (int c2 = 0; c2 < D / 64; ++c2)
#pragma unroll
            for (int u8 = 0; u8 < 8; ++u8)
              asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + 32 * c2 + 4 * u8),
                           "f"(__uint_as_float(dq[c2][4 * u8])), "f"(__uint_as_float(dq[c2][4 * u8 + 1])),
                           "f"(__uint_as_float(dq[c2][4 * u8 + 2])), "f"(__uint_as_float(dq[c2][4 * u8 + 3]))
                           : "memory");
        }
#endif
      }
#elif !defined(FA_EXP_NODRAIN)   // (FA_EXP_NODRAIN: timing-only experiment, dQ is read out of TMEM and dropped)
#pragma unroll
      for (int c2 = 0; c2 < D / 64; ++c2) {
        // use number u of staging box hh by THIS warpgroup; group 0's very first use finds the box free
        const int u = (it >> 1) * (D / 64) + c2;
        mbar_wait(&stg_empty[2 * hh + g], g == 0 ? ((u & 1) ^ 1) : (u & 1));
        uint8_t* dst = stg_row + hh * Cfg::STG_BYTES;
#pragma unroll
        for (int u8 = 0; u8 < 8; ++u8)
          *reinterpret_cast<uint4*>(dst + ((u8 ^ jx) << 4)) =
              make_uint4(dq[c2][4 * u8], dq[c2][4 * u8 + 1], dq[c2][4 * u8 + 2], dq[c2][4 * u8 + 3]);
        fence_proxy_async_smem();
        mbar_arrive(&stg_full[hh]);
      }
#else
      asm volatile("" ::"r"(dq[0][0]), "r"(dq[D / 64 - 1][31]));
#endif
      FA_TR(15)
    }

    // ---- epilogue: group 0 stores dK (scaled), group 1 stores dV; each half stores D/2 columns
    mbar_wait(dkv_done, 0);
    tc_fence_after();
    const uint32_t tacc = lane_base + (g == 0 ? Cfg::T_DK : Cfg::T_DV) + (D / 2) * hh;
    const float mul = (g == 0) ? p.scale : 1.0f;
    __nv_bfloat16* orow = (g == 0 ? dKb : dVb) + static_cast<long long>(k0 + j) * p.sn + (D / 2) * hh;
#pragma unroll
    for (int c = 0; c < D / 64; ++c) {
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
          if ((D / 2) * hh + 32 * c + 8 * i < p.d) *reinterpret_cast<uint4*>(orow + 32 * c + 8 * i) = o;
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 17) tmem_dealloc<512>(tmem_base);
}

}  // namespace sm100
}  // namespace fa
