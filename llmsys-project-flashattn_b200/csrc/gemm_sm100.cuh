// bf16 GEMM on tcgen05 for the Linear layers around the attention core:
//     C[M, N] (fp32 or bf16, row-major) = A[M, K] . B[K, N]       bf16 operands, fp32 accumulation in TMEM
// The reference runs every projection through src/combine.cu:148-210 (one thread per output element, fp32 FMA).
//
// Either operand may be given in either memory order, so the three GEMMs of a Linear layer need no transposes:
//     forward   y  = x . W           A = x  [M][K] (K contiguous: "K-major"),   B = W  [K][N] (N contiguous: "MN-major")
//     backward  dx = dy . W^T        A = dy [M][K'] K-major,                    B = W^T given as W [N'][K'] -> K-major
//               dW = x^T . dy        A = x^T given as x [K'][M'] -> MN-major,   B = dy [K'][N] MN-major
// Persistent CTAs (one per SM) walk 128 x BN output tiles; two accumulator stages in TMEM overlap a tile's epilogue with
// the next tile's MMAs; K consumed 64 elements per stage through a 4-deep TMA ring (128-byte swizzled
// tiles, the same descriptor conventions as the attention kernels: a K-major tile is [128 rows][64 k] with 32-byte
// K steps, an MN-major tile is two [64 k][64 cols] chunks with 2048-byte K steps), tcgen05.mma issued by one lane
// of warp 5, accumulator (128 lanes x 128 fp32 columns) read back by the four epilogue warps and stored row-wise.
// Up to three outputs: the fused Q/K/V projection multiplies x by the concatenated [E][3E] weight once and writes the
// three column blocks to separate (B*N, E) buffers -- the (B, N, nh, d) layout the flash kernels consume in place.
#pragma once
#include "ptx.cuh"

namespace fa {
namespace gemm {

constexpr int BM = 128, BK = 64, NSTAGE = 4;
constexpr int A_BYTES = BM * BK * 2;       // 16 KiB
constexpr int NTHREADS = 192;
// BN = 256 halves the A-operand re-reads per output column (an M128xN128xK16 SS MMA fetches 8 KB of shared memory per
// 64 clk = the whole 128 B/clk of the SM; N = 256 needs 12 KB per 128 clk), BN = 128 is kept for narrow outputs.
template <int BN>
struct Cfg {
  static constexpr int B_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int EPI_BYTES = 4 * 32 * 36 * 4;   // per epilogue warp a [32 rows][36] fp32 transpose tile
  static constexpr int SMEM_BYTES = NSTAGE * STAGE_BYTES + EPI_BYTES + 1024 /*align*/ + 256 /*barriers*/;
};

struct Params {
  int M, N, K;
  void* out[3];          // column block j of width n_split goes to out[j] (n_split == 0: everything to out[0])
  int n_split;
  long long ldo;         // row stride of the outputs (elements)
  int out_bf16;
};

__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}

// c ? a : b on 64-bit values, opaque to the optimiser (which otherwise turns a chain of selects over kernel parameters
// back into an indexed load from a stack copy of them)
__device__ __forceinline__ uint64_t sel64(uint64_t b, uint64_t a, bool c) {
  uint64_t r;
  asm("{\n\t.reg .pred p;\n\tsetp.ne.s32 p, %3, 0;\n\tselp.b64 %0, %1, %2, p;\n\t}" : "=l"(r) : "l"(a), "l"(b), "r"(static_cast<int>(c)));
  return r;
}

template <bool A_MN, bool B_MN, int BN>
__global__ void __launch_bounds__(NTHREADS, 1)
    gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const Params p) {
  constexpr int STAGE_BYTES = Cfg<BN>::STAGE_BYTES;
  constexpr int ACC_STAGES = 512 / BN >= 2 ? 2 : 1;      // accumulator stages in TMEM (2 x 256 or 2 x 128 columns)
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  float* epi = reinterpret_cast<float*>(smem + NSTAGE * STAGE_BYTES);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + NSTAGE * STAGE_BYTES + Cfg<BN>::EPI_BYTES);
  uint64_t* full = bars;                 // [NSTAGE]
  uint64_t* empty = bars + NSTAGE;       // [NSTAGE]
  uint64_t* acc_full = bars + 2 * NSTAGE;      // [2]  accumulator stage complete (MMA -> epilogue)
  uint64_t* acc_empty = bars + 2 * NSTAGE + 2; // [2]  accumulator stage drained  (epilogue -> MMA)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * NSTAGE + 4);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nk = (p.K + BK - 1) / BK;
  // PERSISTENT: a CTA walks output tiles t = blockIdx.x, + gridDim.x, ... (grouped order, see tile_origin).  Two
  // accumulator stages in TMEM let the epilogue of tile i (TMEM -> registers -> global) run under the MMAs of tile i+1.
  const int tiles_n = (p.N + BN - 1) / BN, tiles_m = (p.M + BM - 1) / BM;
  const int n_tiles = tiles_n * tiles_m;
  // tile t -> (m0, n0): bands of GROUP_M row panels, M index fastest inside a band, so the ~148 tiles in flight cover GROUP_M row panels x ~148/GROUP_M column panels -- a squarer footprint than one
  // row of tiles (8192^3: B was re-streamed from HBM once per 4.6 row panels, 2.0 GB of DRAM reads per launch)
  constexpr int GROUP_M = 8;
  auto tile_origin = [&](int t, int& m0, int& n0) {
    const int band = t / (GROUP_M * tiles_n);
    const int first_m = band * GROUP_M;
    const int rows = min(GROUP_M, tiles_m - first_m);
    const int r = t - band * GROUP_M * tiles_n;
    m0 = (first_m + r % rows) * BM;
    n0 = (r / rows) * BN;
  };

  if (warp == 4 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int i = 0; i < NSTAGE; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&acc_full[i], 1);
      mbar_init(&acc_empty[i], 128);
    }
    fence_mbar_init();
  }
  if (warp == 5) tmem_alloc<ACC_STAGES * BN>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 4) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      uint32_t it = 0;     // stage counter across tiles
      for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        int m0, n0;
        tile_origin(t, m0, n0);
        for (int kt = 0; kt < nk; ++kt, ++it) {
          const int s = it % NSTAGE;
          mbar_wait(&empty[s], ((it / NSTAGE) & 1) ^ 1);
          mbar_expect_tx(&full[s], STAGE_BYTES);
          uint8_t* sa = smem + s * STAGE_BYTES;
          uint8_t* sb = sa + A_BYTES;
          const int k0 = kt * BK;
          if (A_MN) {       // memory [K][M]: two [64 k][64 m] chunks
            tma_load_2d(sa, &tmA, &full[s], m0, k0);
            tma_load_2d(sa + A_BYTES / 2, &tmA, &full[s], m0 + 64, k0);
          } else {          // memory [M][K]: one [128 m][64 k] tile
            tma_load_2d(sa, &tmA, &full[s], k0, m0);
          }
          if (B_MN) {       // memory [K][N]: BN / 64 chunks of [64 k][64 n]
#pragma unroll
            for (int c = 0; c < BN / 64; ++c) tma_load_2d(sb + c * 8192, &tmB, &full[s], n0 + 64 * c, k0);
          } else {          // memory [N][K]: one [BN n][64 k] tile
            tma_load_2d(sb, &tmB, &full[s], k0, n0);
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 5) {
    // ------------------------------------------------------------------ MMA issuer (uniform control flow, one lane issues)
    const bool leader = elect_one();
    constexpr uint32_t idesc = make_idesc_bf16(BM, BN, A_MN ? 1 : 0, B_MN ? 1 : 0);
    const uint64_t da = A_MN ? make_smem_desc(smem_u32(smem), A_BYTES / 2, 1024) : make_smem_desc(smem_u32(smem), 16, 1024);
    const uint64_t db = B_MN ? make_smem_desc(smem_u32(smem + A_BYTES), 8192, 1024)
                             : make_smem_desc(smem_u32(smem + A_BYTES), 16, 1024);
    const uint32_t a_lo = static_cast<uint32_t>(da), a_hi = static_cast<uint32_t>(da >> 32);
    const uint32_t b_lo = static_cast<uint32_t>(db), b_hi = static_cast<uint32_t>(db >> 32);
    constexpr uint32_t a_step = (A_MN ? 2048 : 32) >> 4, b_step = (B_MN ? 2048 : 32) >> 4;   // per 16 k elements
    uint32_t it = 0, tc = 0;
    for (int t = blockIdx.x; t < n_tiles; t += gridDim.x, ++tc) {
      const uint32_t as = tc % ACC_STAGES;
      if (tc >= ACC_STAGES) {      // the epilogue has drained this accumulator stage (its previous tile)
        mbar_wait(&acc_empty[as], ((tc / ACC_STAGES) - 1) & 1);
        tc_fence_after();
      }
      const uint32_t tacc = tmem_base + as * BN;
      for (int kt = 0; kt < nk; ++kt, ++it) {
        const int s = it % NSTAGE;
        mbar_wait(&full[s], (it / NSTAGE) & 1);
        tc_fence_after();
        if (leader) {
          const uint32_t so = (s * STAGE_BYTES) >> 4;
#pragma unroll
          for (int k = 0; k < BK / 16; ++k)
            mma_ss2(tacc, a_lo + so + k * a_step, a_hi, b_lo + so + k * b_step, b_hi, idesc, (kt > 0 || k > 0) ? 1u : 0u);
          mma_commit(&empty[s]);
        }
        __syncwarp();
      }
      if (leader) mma_commit(&acc_full[as]);
      __syncwarp();
    }
  } else {
    // ------------------------------------------------------------------ epilogue: TMEM -> registers -> global
    uint32_t tc = 0;
    for (int t = blockIdx.x; t < n_tiles; t += gridDim.x, ++tc) {
      int m0, n0;
      tile_origin(t, m0, n0);
      const uint32_t as = tc % ACC_STAGES;
      mbar_wait(&acc_full[as], (tc / ACC_STAGES) & 1);
      tc_fence_after();
      const int row = m0 + warp * 32 + lane;
      const uint32_t taddr = tmem_base + as * BN + (static_cast<uint32_t>(warp * 32) << 16);
#pragma unroll 1
      for (int c = 0; c < BN / 32; ++c) {
        uint32_t u[32];
        tmem_ld32(taddr + 32 * c, u);
        tmem_wait_ld();
        if (c == BN / 32 - 1) {      // the whole stage is in registers (or already stored): hand it back to the MMA warp
          tc_fence_before();
          mbar_arrive(&acc_empty[as]);
        }
        const int col0 = n0 + 32 * c;
        // destination: column block j of width n_split lives in out[j]
        int j = 0, cj = col0;
        if (p.n_split > 0) {
          j = col0 / p.n_split;
          cj = col0 - j * p.n_split;
        }
        // (selected with constant indices: `p.out[j]` with a run-time j makes the compiler copy the kernel parameters to
        // a local-memory stack frame and re-load the pointer from it for every store -- ncu showed the epilogue warps
        // stalled ~300 clk per store group on exactly those loads, 2600 clk per 32-column piece)
        char* const outp = reinterpret_cast<char*>(
            sel64(sel64(reinterpret_cast<uint64_t>(p.out[2]), reinterpret_cast<uint64_t>(p.out[1]), j == 1),
                  reinterpret_cast<uint64_t>(p.out[0]), j == 0));
        const int valid = min(32, p.N - col0);      // (<= 0 when the piece is past the last column)
        if ((p.ldo & 3) == 0 && (cj & 3) == 0 && (valid & 3) == 0) {
          // Coalesced path: the accumulator arrives with lane = row; a per-warp shared-memory transpose turns it into
          // lane = 4 consecutive columns, so one store instruction writes 4 rows x 128 contiguous bytes instead of 32
          // scattered 16-byte pieces (the output of a short-K GEMM is a gigabyte: 131072 x 2048 fp32).
          float* T = epi + warp * (32 * 36);
#pragma unroll
          for (int k = 0; k < 8; ++k)
            *reinterpret_cast<float4*>(T + lane * 36 + 4 * k) =
                make_float4(__uint_as_float(u[4 * k]), __uint_as_float(u[4 * k + 1]), __uint_as_float(u[4 * k + 2]),
                            __uint_as_float(u[4 * k + 3]));
          __syncwarp();
          const int cq = (lane & 7) * 4;      // my 4 columns of the piece
#pragma unroll
          for (int rr = 0; rr < 32; rr += 4) {
            const int r = rr + (lane >> 3);
            const float4 v = *reinterpret_cast<const float4*>(T + r * 36 + cq);
            const int grow = m0 + warp * 32 + r;
            if (grow < p.M && cq < valid) {
              if (p.out_bf16) {
                __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(outp) + static_cast<long long>(grow) * p.ldo + cj + cq;
                *reinterpret_cast<uint2*>(dst) = make_uint2(pack_bf16x2(v.x, v.y), pack_bf16x2(v.z, v.w));
              } else {
                float* dst = reinterpret_cast<float*>(outp) + static_cast<long long>(grow) * p.ldo + cj + cq;
                *reinterpret_cast<float4*>(dst) = v;
              }
            }
          }
          __syncwarp();      // the tile is reused by the next piece
        } else if (row < p.M && valid > 0) {
          // ragged / unaligned outputs: element-wise stores, lane = row
          if (p.out_bf16) {
            __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(outp) + static_cast<long long>(row) * p.ldo + cj;
#pragma unroll
            for (int i = 0; i < 32; ++i)      // (predicated, fully unrolled: a run-time index would spill u[] to local memory)
              if (i < valid) dst[i] = __float2bfloat16_rn(__uint_as_float(u[i]));
          } else {
            float* dst = reinterpret_cast<float*>(outp) + static_cast<long long>(row) * p.ldo + cj;
#pragma unroll
            for (int i = 0; i < 32; ++i)
              if (i < valid) dst[i] = __uint_as_float(u[i]);
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 5) tmem_dealloc<ACC_STAGES * BN>(tmem_base);
}

}  // namespace gemm
}  // namespace fa
