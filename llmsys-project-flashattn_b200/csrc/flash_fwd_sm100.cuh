// Flash-attention forward for sm_100a: TMA -> shared memory -> tcgen05.mma -> TMEM.
//
// One CTA owns 256 query rows of one (batch, head): two 128-row Q tiles that ping-pong so the
// tensor pipe works on one tile while the other tile's softmax runs on the CUDA cores.
//   warps 0-15 : four softmax groups.  Group (g, hh) = Q tile g, column half hh: thread t of warp w
//                owns query row 32w+t of tile g (= its TMEM lane) and 2 x 32 of the 128 score columns.
//                Two threads per row give every SM sub-partition four softmax warps to interleave
//                (with one thread per row the phase was latency-bound at ~5 clk per instruction);
//                the two halves agree on the row max through a 4 KB shared-memory exchange.
//   warp 16    : TMA producer (Q tiles once; K,V tiles through an NSTAGE ring)
//   warp 17    : tcgen05.mma issuer (one lane) + TMEM allocation
//   warps 18-19: idle (they only complete the helper warpgroup so that setmaxnreg can move registers:
//                20 warps launch with 96 registers each; the softmax warpgroups grow to 104, the helper
//                warpgroup shrinks to 64)
// TMEM (512 columns): S0 | S1 (128 fp32 columns each), O0 | O1 (D columns each).  P (bf16)
// overwrites the first 64 columns of its S tile and is consumed directly from TMEM as the
// A operand of the PV MMA, so P never touches shared memory.
// Per KV tile j and Q tile g the issuer runs   O_g += P_g(j) V_j ;  S_g = Q_g K_{j+1}^T
// and the softmax groups of g turn S_g into P_g: row max, lazy rescale of O_g (only when the max
// grew by more than 2^8), exp2 with the 1/sqrt(d)*log2(e) scale folded into one FFMA, row sum.
// Masks: causal (tiles above the diagonal are skipped, only the diagonal tile is masked),
// key padding as kv_len[b] (tiles beyond it are skipped) or a generic additive (B,N) mask.
#pragma once
#include "ptx.cuh"

namespace fa {
namespace sm100 {

struct FwdParams {
  int B, H, N;
  int d;                  // real head dim (<= D): columns d..D-1 of every tile are TMA zero-fill and are never stored
  const int* kv_len;      // device int32[B] or nullptr
  const float* key_mask;  // device fp32 (B,N) or nullptr (MASKMODE 2)
  void* O;                // (B,H,N,D) OutT with strides below
  long long o_sb, o_sh, o_sn;
  float* M;               // (B,H,N) row max of scaled scores
  float* L;               // (B,H,N) sum exp(s - m)
  float scale;            // 1/sqrt(d)
  float scale_log2;       // scale * log2(e)
  long long* trace;       // bring-up only (FA_TRACE builds)
  // persistent variant only (flash_fwd_persistent_sm100.cuh)
  int n_qblk;             // ceil(N / 256)
  int n_items;            // n_qblk * H * B
  int qb_major;           // item order: 0 = query blocks of one head adjacent, 1 = all heads' longest blocks first
};

#ifdef FA_TRACE
#define FA_FTR(slot) \
  if (tr && j < 48) tr[j * 32 + (slot)] = clock64();
#else
#define FA_FTR(slot)
#endif

template <int D>
struct FwdCfg {
  static constexpr int NCHUNK = D / 64;            // 128-byte swizzle chunks per row
  static constexpr int CHUNK_BYTES = 128 * 128;    // [128 rows][64 bf16]
  static constexpr int TILE_BYTES = NCHUNK * CHUNK_BYTES;
  static constexpr int NSTAGE = (D == 128) ? 4 : 8;
  static constexpr int SMEM_TILES = 2 * TILE_BYTES + NSTAGE * TILE_BYTES;
  static constexpr int XCHG_BYTES = 2 * 2 * 2 * 128 * 4;  // [parity][tile][half][row] fp32 row-max / row-sum exchange
  static constexpr int SMEM_BYTES = SMEM_TILES + XCHG_BYTES + 1024 /*align*/ + 256 /*barriers*/;
  static constexpr int S_COL0 = 0, S_COL1 = 128, O_COL0 = 256, O_COL1 = 256 + D;
  static constexpr int NTHREADS = 640;
};


template <typename OutT>
__device__ __forceinline__ void store_row32(OutT* dst, const float* v);
template <>
__device__ __forceinline__ void store_row32<float>(float* dst, const float* v) {
#pragma unroll
  for (int i = 0; i < 8; ++i)
    reinterpret_cast<float4*>(dst)[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
}
template <>
__device__ __forceinline__ void store_row32<__nv_bfloat16>(__nv_bfloat16* dst, const float* v) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    uint4 u;
    u.x = pack_bf16x2(v[8 * i], v[8 * i + 1]);
    u.y = pack_bf16x2(v[8 * i + 2], v[8 * i + 3]);
    u.z = pack_bf16x2(v[8 * i + 4], v[8 * i + 5]);
    u.w = pack_bf16x2(v[8 * i + 6], v[8 * i + 7]);
    reinterpret_cast<uint4*>(dst)[i] = u;
  }
}

// 32 output columns of one row, of which the first `valid` (a multiple of 8, possibly 0) exist (head dims below D)
template <typename OutT>
__device__ __forceinline__ void store_row32_upto(OutT* dst, const float* v, int valid) {
  if (valid >= 32) {
    store_row32<OutT>(dst, v);
  } else {
    // (fully unrolled and predicated: a run-time index would force the caller's register array into local memory --
    // for the full-width path too, since both read the same array)
#pragma unroll
    for (int i = 0; i < 32; ++i)
      if (i < valid) dst[i] = static_cast<OutT>(v[i]);
  }
}

// scheduling fence: everything that produces these registers is ordered before the next volatile asm
__device__ __forceinline__ void pin16(const uint32_t (&r)[16], float f) {
  asm volatile("" ::"r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
               "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "f"(f));
}

#ifndef FA_FWD_EMU
#define FA_FWD_EMU 1   // column PAIRS of every 8 scores whose exponentials are evaluated by the packed polynomial on the
                       // FMA pipe (ex2_poly2) instead of MUFU.EX2.  Interleaved in-process A/B on B200 at cfg4, sustained
                       // (power-capped) clocks, tools/ab_kernels.py: 0 -> 1.496 ms, 1 -> 1.39-1.40 ms, 2 -> 1.56 ms.
                       // (An earlier scalar variant, 9 issue slots per exponential, was slower at every setting.)
#endif
// MASKMODE: 0 none (N-ragged only), 1 kv_len[b], 2 additive key mask (B,N)
template <int D, bool CAUSAL, int MASKMODE, typename OutT>
__global__ void __launch_bounds__(640, 1)
    fwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
               const __grid_constant__ CUtensorMap tmV, const FwdParams p) {
  using Cfg = FwdCfg<D>;
  constexpr int NSTAGE = Cfg::NSTAGE;
#ifndef FA_FWD_EMU_D64
#define FA_FWD_EMU_D64 FA_FWD_EMU
#endif
  constexpr int EMU = (D == 64) ? FA_FWD_EMU_D64 : FA_FWD_EMU;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // align by offset (not by pointer cast) so the compiler keeps the shared address space
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sQ = smem;                                   // [2][TILE_BYTES]
  uint8_t* sKV = smem + 2 * Cfg::TILE_BYTES;            // [NSTAGE][TILE_BYTES]
  float* xchg = reinterpret_cast<float*>(smem + Cfg::SMEM_TILES);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::SMEM_TILES + Cfg::XCHG_BYTES);
  uint64_t* q_full = bars;                 // [2]
  uint64_t* kv_full = bars + 2;            // [NSTAGE]
  uint64_t* kv_empty = kv_full + NSTAGE;   // [NSTAGE]
  uint64_t* s_full = kv_empty + NSTAGE;    // [2]
  uint64_t* p_lo = s_full + 2;             // [2]  P of keys [0,64) of the tile written
  uint64_t* p_hi = p_lo + 2;               // [2]  P of keys [64,128) written
  uint64_t* o_done = p_hi + 2;             // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_done + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#ifdef FA_TRACE
  long long* tr = (blockIdx.x == 1 && blockIdx.y == 0 && blockIdx.z == 0 && lane == 0) ? p.trace : nullptr;
#endif
  const int qb = CAUSAL ? (gridDim.x - 1 - blockIdx.x) : blockIdx.x;  // causal: longest blocks first
  const int h = blockIdx.y, b = blockIdx.z;
  int kv_end = p.N;
  if (MASKMODE == 1) kv_end = min(kv_end, max(__ldg(p.kv_len + b), 0));
  const int nkv_total = (kv_end + 127) >> 7;
  int r0[2], nkv[2];
#pragma unroll
  for (int g = 0; g < 2; ++g) {
    r0[g] = qb * 256 + g * 128;
    nkv[g] = (r0[g] >= p.N) ? 0 : (CAUSAL ? min(nkv_total, (r0[g] >> 7) + 1) : nkv_total);
  }
  const int nk = max(nkv[0], nkv[1]);

  if (warp == 16 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&q_full[i], 1);
      mbar_init(&s_full[i], 1);
      mbar_init(&p_lo[i], 256);
      mbar_init(&p_hi[i], 256);
      mbar_init(&o_done[i], 1);
    }
    for (int i = 0; i < NSTAGE; ++i) {
      mbar_init(&kv_full[i], 1);
      mbar_init(&kv_empty[i], 1);
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
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
#pragma unroll
      for (int g = 0; g < 2; ++g) {
        if (nkv[g] > 0) {
          mbar_expect_tx(&q_full[g], Cfg::TILE_BYTES);
#pragma unroll
          for (int c = 0; c < Cfg::NCHUNK; ++c)
            tma_load_4d(sQ + g * Cfg::TILE_BYTES + c * Cfg::CHUNK_BYTES, &tmQ, &q_full[g], c * 64, r0[g], h, b);
        }
      }
      for (int t = 0; t < 2 * nk; ++t) {
        const int stage = t % NSTAGE, it = t / NSTAGE;
        mbar_wait(&kv_empty[stage], (it & 1) ^ 1);
        mbar_expect_tx(&kv_full[stage], Cfg::TILE_BYTES);
        const CUtensorMap* tm = (t & 1) ? &tmV : &tmK;
        uint8_t* dst = sKV + stage * Cfg::TILE_BYTES;
#pragma unroll
        for (int c = 0; c < Cfg::NCHUNK; ++c)
          tma_load_4d(dst + c * Cfg::CHUNK_BYTES, tm, &kv_full[stage], c * 64, (t >> 1) * 128, h, b);
      }
    }
    __syncwarp();
  } else if (warp == 17) {
    // ------------------------------------------------------------------ MMA issuer
    // The whole warp walks the schedule (uniform control flow, so the descriptors live in uniform
    // registers and no per-lane serialisation loop is generated); one elected lane issues.
    if (nk > 0) {
      const bool leader = elect_one();
      constexpr uint32_t idesc_qk = make_idesc_bf16(128, 128, 0, 0);
      constexpr uint32_t idesc_pv = make_idesc_bf16(128, D, 0, 1);
      const uint32_t tS[2] = {tmem_base + Cfg::S_COL0, tmem_base + Cfg::S_COL1};
      const uint32_t tO[2] = {tmem_base + Cfg::O_COL0, tmem_base + Cfg::O_COL1};
      // descriptor = {low word: (address >> 4) | LBO field, high word: SBO | version | swizzle}
      const uint64_t dq = make_smem_desc(smem_u32(sQ), 16, 1024);
      const uint64_t dk = make_smem_desc(smem_u32(sKV), 16, 1024);
      const uint64_t dv = make_smem_desc(smem_u32(sKV), Cfg::CHUNK_BYTES, 1024);
      const uint32_t q_lo = static_cast<uint32_t>(dq), k_lo = static_cast<uint32_t>(dk),
                     v_lo = static_cast<uint32_t>(dv);
      const uint32_t kq_hi = static_cast<uint32_t>(dk >> 32), v_hi = static_cast<uint32_t>(dv >> 32);
      auto wait_full = [&](int t) {
        mbar_wait(&kv_full[t % NSTAGE], (t / NSTAGE) & 1);
        tc_fence_after();
      };
      // S_g = Q_g K^T : A = Q (K-major), B = K tile (K-major), 16 head-dim elements per MMA
      auto issue_qk = [&](int g, int t) {
        const uint32_t qa = q_lo + g * (Cfg::TILE_BYTES >> 4), ka = k_lo + (t % NSTAGE) * (Cfg::TILE_BYTES >> 4);
        if (leader) {
#pragma unroll
          for (int k = 0; k < D / 16; ++k) {
            const uint32_t off = ((k >> 2) * Cfg::CHUNK_BYTES + (k & 3) * 32) >> 4;
            mma_ss2(tS[g], qa + off, kq_hi, ka + off, kq_hi, idesc_qk, k > 0);
          }
        }
      };
      // O_g += P_g V : A = P (bf16 in TMEM, 8 columns per 16 keys), B = V tile (MN-major), 16 keys per MMA
      // (issued in two halves of 64 keys so the first half overlaps the second half of the softmax)
      auto issue_pv = [&](int g, int t, bool acc, int half) {
        const uint32_t va = v_lo + (t % NSTAGE) * (Cfg::TILE_BYTES >> 4);
        if (leader) {
#pragma unroll
          for (int k = 4 * half; k < 4 * half + 4; ++k)
            mma_ts2(tO[g], tS[g] + k * 8, va + k * (2048 >> 4), v_hi, idesc_pv, (acc || k > 0) ? 1u : 0u);
        }
      };
      auto commit = [&](uint64_t* bar) {
        if (leader) mma_commit(bar);
      };
      wait_full(0);
#pragma unroll
      for (int g = 0; g < 2; ++g) {
        if (nkv[g] > 0) {
          mbar_wait(&q_full[g], 0);
          tc_fence_after();
          issue_qk(g, 0);
          commit(&s_full[g]);
        }
      }
      commit(&kv_empty[0]);
      for (int j = 0; j < nk; ++j) {
        const int tv = 2 * j + 1, tk = 2 * j + 2;
        wait_full(tv);
        bool k_ready = false;
#pragma unroll
        for (int g = 0; g < 2; ++g) {
          if (j < nkv[g]) {
            mbar_wait(&p_lo[g], j & 1);
            tc_fence_after();
            FA_FTR(0 + 4 * g)
            issue_pv(g, tv, j > 0, 0);
            mbar_wait(&p_hi[g], j & 1);
            tc_fence_after();
            issue_pv(g, tv, j > 0, 1);
            commit(&o_done[g]);
            FA_FTR(1 + 4 * g)
            if (j + 1 < nkv[g]) {
              if (!k_ready) {
                wait_full(tk);
                k_ready = true;
              }
              FA_FTR(2 + 4 * g)
              issue_qk(g, tk);
              commit(&s_full[g]);
              FA_FTR(3 + 4 * g)
            }
          }
        }
        commit(&kv_empty[tv % NSTAGE]);
        if (j + 1 < nk) commit(&kv_empty[tk % NSTAGE]);
      }
    }
    __syncwarp();
   }
  } else {
    // ------------------------------------------------------------------ softmax groups
    reg_alloc<104>();
    const int wg = warp >> 2, w = warp & 3;
    const int g = wg & 1, hh = wg >> 1;          // Q tile, column half
    const int rl = w * 32 + lane;                // row inside the tile == TMEM lane
    // The two warps that share 32 rows (column halves hh = 0 / 1) synchronise on their own named barrier:
    // the rescale decision below is taken per warp, so anything wider could deadlock on divergent data.
    const uint32_t pair_bar = 1 + g * 4 + w;
    const int row = r0[g] + rl;
    const uint32_t lane_base = tmem_base + (static_cast<uint32_t>(w * 32) << 16);
    const uint32_t tSg = lane_base + (g ? Cfg::S_COL1 : Cfg::S_COL0);
    // A thread owns 2 x 32 score columns of its row: chunk A = keys [32hh, 32hh+32) and chunk B = keys
    // [64+32hh, 96+32hh).  Both are loaded up front (64 live scores); the chunk-A exponentials of the two
    // halves complete keys [0,64) of P, which is signalled separately (p_lo) so that half of P.V runs on
    // the tensor pipe while the chunk-B exponentials are still being evaluated.
    const uint32_t tA = tSg + 32 * hh, tB = tSg + 64 + 32 * hh;
    const uint32_t tPA = tSg + 16 * hh, tPB = tSg + 32 + 16 * hh;   // packed P columns of chunk A / B
    const uint32_t tO = lane_base + (g ? Cfg::O_COL1 : Cfg::O_COL0) + (D / 2) * hh;
    const float sc = (MASKMODE == 2) ? 1.0f : p.scale_log2;
    constexpr float LOG2E = 1.4426950408889634f;
    const float* mrow = (MASKMODE == 2) ? p.key_mask + static_cast<long long>(b) * p.N : nullptr;
    OutT* orow = reinterpret_cast<OutT*>(p.O) + b * p.o_sb + h * p.o_sh + static_cast<long long>(row) * p.o_sn +
                 (D / 2) * hh;
    const long long stat_idx = (static_cast<long long>(b) * p.H + h) * p.N + row;
    float* x_mine = xchg + (g * 2 + hh) * 128 + rl;        // + parity * 512
    float* x_peer = xchg + (g * 2 + (hh ^ 1)) * 128 + rl;

    if (nkv[g] == 0) {
      if (row < p.N) {  // no visible key at all (kv_len == 0): O = 0, m = -inf, l = 0
        float z[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) z[i] = 0.f;
#pragma unroll
        for (int c = 0; c < D / 64; ++c) store_row32_upto<OutT>(orow + 32 * c, z, p.d - (D / 2) * hh - 32 * c);
        if (hh == 0) {
          p.M[stat_idx] = -INFINITY;
          p.L[stat_idx] = 0.f;
        }
      }
    } else {
      float m_used = -INFINITY, m_true = -INFINITY, l_run = 0.f;
      // masks for the 32 scores of my row starting at key `key0` (generic additive mask, ragged / padded
      // keys, causal diagonal tile)
      auto mask_chunk = [&](int key0, float(&s)[32]) {
        if (MASKMODE == 2) {
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const float mv = (key0 + i < p.N) ? __ldg(mrow + key0 + i) : 0.f;
            s[i] = fmaf(s[i], p.scale_log2, mv * LOG2E);
          }
        }
        if ((key0 + 32 > kv_end) || (CAUSAL && (key0 + 31 > r0[g]))) {
          int limit = kv_end - key0;
          if (CAUSAL) limit = min(limit, row - key0 + 1);
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (i >= limit) s[i] = -INFINITY;
        }
      };
      auto max32 = [](const float(&s)[32]) {   // 16 FMNMX3 in four chains
        float a0 = fmax3(s[0], s[1], s[2]), a1 = fmax3(s[3], s[4], s[5]);
        float a2 = fmax3(s[6], s[7], s[8]), a3 = fmax3(s[9], s[10], s[11]);
#pragma unroll
        for (int i = 12; i < 28; i += 8) {
          a0 = fmax3(a0, s[i], s[i + 1]), a1 = fmax3(a1, s[i + 2], s[i + 3]);
          a2 = fmax3(a2, s[i + 4], s[i + 5]), a3 = fmax3(a3, s[i + 6], s[i + 7]);
        }
        return fmaxf(fmax3(a0, a1, s[28]), fmax3(a2, a3, fmax3(s[29], s[30], s[31])));
      };
      // exp2 of 32 scores -> 16 packed bf16x2 registers, returns their sum.  x = s*sc - m*sc and the row sum
      // are evaluated two columns per instruction (FFMA2 / FADD2).  Of every 8 exponentials EMU are evaluated
      // by ex2_poly on the FMA pipe, the rest by MUFU.EX2 (see ptx.cuh).
      auto exp_pack = [&](const float(&s)[32], float neg_m, uint32_t(&pk)[16]) {
        const uint64_t sc2 = f32x2(sc, sc), nm2 = f32x2(neg_m, neg_m);
        uint64_t ra = f32x2(0.f, 0.f), rb = ra;
#pragma unroll
        for (int i = 0; i < 32; i += 8) {
          float e[8];
#pragma unroll
          for (int t = 0; t < 8; t += 2) {
            float x0, x1;
            f32x2_unpack(fma_f32x2(f32x2(s[i + t], s[i + t + 1]), sc2, nm2), x0, x1);
            // EMU of the four column PAIRS of every 8 go to the packed polynomial (FMA pipe), the rest to MUFU.EX2
#ifdef FA_FWD_EMU_MASK     // experiment: bit k set = column pair k of the 16 pairs of a 32-score chunk uses the polynomial
            const bool emu = ((FA_FWD_EMU_MASK >> ((i + t) / 2)) & 1) != 0;
#else
            const bool emu = (EMU >= 1 && t == 6) || (EMU >= 2 && t == 2) || (EMU >= 3 && t == 4) || (EMU >= 4 && t == 0);
#endif
            if (emu) {
              ex2_poly2(x0, x1, e[t], e[t + 1]);
            } else {
              e[t] = ex2_approx(x0);
              e[t + 1] = ex2_approx(x1);
            }
          }
          ra = add_f32x2(ra, add_f32x2(f32x2(e[0], e[1]), f32x2(e[4], e[5])));
          rb = add_f32x2(rb, add_f32x2(f32x2(e[2], e[3]), f32x2(e[6], e[7])));
#pragma unroll
          for (int t = 0; t < 4; ++t) pk[i / 2 + t] = pack_bf16x2(e[2 * t], e[2 * t + 1]);
        }
        float s0, s1;
        f32x2_unpack(add_f32x2(ra, rb), s0, s1);
        return s0 + s1;
      };
      for (int j = 0; j < nkv[g]; ++j) {
        mbar_wait(&s_full[g], j & 1);
        tc_fence_after();
        FA_FTR(8 + 4 * wg)
        const int kA = j * 128 + 32 * hh, kB = kA + 64;
        float sA[32], sB[32];
        tmem_ld32f(tA, sA);      // both chunks in flight, one wait: 64 scores live
        tmem_ld32f(tB, sB);
        tmem_wait_ld();
        mask_chunk(kA, sA);
        mask_chunk(kB, sB);
        float mx = fmaxf(max32(sA), max32(sB));
        x_mine[(j & 1) * 512] = mx;
        FA_FTR(9 + 4 * wg)
        // Speculate that the reference maximum m_used survives this tile (it does unless some row maximum
        // grows by more than 2^8): the chunk-A exponentials start right away on the MUFU pipe, the row-max
        // reduction above shares their shadow on the ALU pipe, and the exchange with the other half of the
        // row happens after them, when the partner has long arrived.  (j == 0 always takes the redo path.)
        float neg_m = -m_used * sc;
        uint32_t pk[16];
        FA_FTR(28 + wg)
        float sumA = exp_pack(sA, neg_m, pk);
        pin16(pk, sumA);              // keep the speculative work ahead of the barrier
        named_bar_sync(pair_bar, 64);   // both halves hold their scores in registers (P may overwrite S) and published
        FA_FTR(10 + 4 * wg)
        mx = fmaxf(mx, x_peer[(j & 1) * 512]);
        m_true = fmaxf(m_true, mx);
        // lazy rescale: only when some row of this warp grew by more than 2^8.  The partner warp of the
        // other half sees the same 32 row maxima, so both take the same decision.
        const bool want = (j == 0) || ((mx - m_used) * sc > 8.0f);
        if (__any_sync(0xffffffffu, want)) {
          if (j == 0) {
            m_used = (mx == -INFINITY) ? 0.f : mx;
          } else {
            const float m_new = fmaxf(m_used, mx);
            const float factor = ex2_approx((m_used - m_new) * sc);
            m_used = m_new;
            l_run *= factor;
            mbar_wait(&o_done[g], (j - 1) & 1);
            tc_fence_after();
#pragma unroll 1
            for (int c = 0; c < D / 16; ++c) {   // my half of the O columns, 8 at a time (rare path)
              uint32_t u[8];
              tmem_ld8(tO + 8 * c, u);
              tmem_wait_ld();
#pragma unroll
              for (int i = 0; i < 8; ++i) u[i] = __float_as_uint(__uint_as_float(u[i]) * factor);
              tmem_st8(tO + 8 * c, u);
            }
          }
          neg_m = -m_used * sc;
          // redo chunk A against the new reference.  Its scores are re-read from TMEM (keeping them live
          // across the speculative pass would cost 32 registers); nobody has stored P yet, and the extra
          // barrier keeps the partner's P store behind this load.
          tmem_ld32f(tA, sA);
          tmem_wait_ld();
          mask_chunk(kA, sA);
          named_bar_sync(pair_bar, 64);
          sumA = exp_pack(sA, neg_m, pk);
        }
        l_run += sumA;
        tmem_st16(tPA, pk);
        tmem_wait_st();
        tc_fence_before();
        mbar_arrive(&p_lo[g]);      // keys [0,64) of P are in TMEM: the first half of P.V may start
        FA_FTR(24 + wg)
        asm volatile("" : "+f"(neg_m));   // chunk B's exponentials stay behind the p_lo signal
        l_run += exp_pack(sB, neg_m, pk);
        tmem_st16(tPB, pk);
        tmem_wait_st();
        tc_fence_before();
        mbar_arrive(&p_hi[g]);
        FA_FTR(11 + 4 * wg)
      }
      // epilogue: combine the two halves' row sums, O / l -> global, statistics
      const int par = nkv[g] & 1;
      x_mine[par * 512] = l_run;
      named_bar_sync(pair_bar, 64);
      const float l_tot = l_run + x_peer[par * 512];
      mbar_wait(&o_done[g], (nkv[g] - 1) & 1);
      tc_fence_after();
      const float inv = (l_tot > 0.f) ? 1.0f / l_tot : 0.f;
#pragma unroll
      for (int c = 0; c < D / 64; ++c) {
        float o[32];
        tmem_ld32f(tO + 32 * c, o);
        tmem_wait_ld();
#pragma unroll
        for (int i = 0; i < 32; ++i) o[i] *= inv;
        if (row < p.N) store_row32_upto<OutT>(orow + 32 * c, o, p.d - (D / 2) * hh - 32 * c);
      }
      if (row < p.N && hh == 0) {
        const float m_out = (MASKMODE == 2) ? m_true * (1.0f / LOG2E) : m_true * p.scale;
        p.M[stat_idx] = m_out;
        p.L[stat_idx] = (m_true == -INFINITY) ? 0.f : l_tot * ex2_approx((m_used - m_true) * sc);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 17) tmem_dealloc<512>(tmem_base);
}

}  // namespace sm100
}  // namespace fa
