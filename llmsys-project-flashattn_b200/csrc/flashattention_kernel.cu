// flashattention_kernel.so -- C ABI of the fused attention path (see include/flashattn_b200.h).
//
// Replaces the reference's src/flashattention_kernel.cu (launchers :259, :352, :694, :761).
// Two arithmetic modes sit behind the same entry points:
//   FA_MODE_BF16 / FA_DTYPE_BF16 : flash_fwd_sm100.cuh / flash_bwd_sm100.cuh -- TMA + tcgen05 + TMEM
//   FA_MODE_FP32 / FA_DTYPE_F32  : flash_fp32.cuh -- fp32 CUDA-core kernels, any N and head dim
// The legacy host-pointer symbols stage through a grow-only device pool (no per-call
// cudaMalloc/cudaFree as in the reference, :280-324) and never exit() the process.
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "host_common.cuh"
#include "tma_host.cuh"
#include "flash_fp32.cuh"
#include "flash_fwd_sm100.cuh"
#include "flash_fwd_persistent_sm100.cuh"
#include "flash_bwd_sm100.cuh"
#include "flash_decode.cuh"

namespace fa {

static int g_mode = -1;  // resolved lazily from MINITORCH_FA_MODE
static long long* g_trace = nullptr;  // FA_TRACE bring-up builds only

// Deterministic bf16 backward: the tensor-core backward accumulates dQ across KV-tile CTAs with fp32 TMA add-reductions
// whose order varies from run to run (results reproducible to fp32 rounding of the sum, far below one bf16 ulp, but not
// bitwise).  With this switch (or env MINITORCH_FA_DETERMINISTIC=1) the bf16 backward runs the two-kernel CUDA-core
// path instead (no atomics, bitwise reproducible, ~40x slower); the forward and the fp32 mode are always deterministic.
static int g_deterministic = -1;
static bool deterministic() {
  if (g_deterministic < 0) {
    const char* e = getenv("MINITORCH_FA_DETERMINISTIC");
    g_deterministic = (e && atoi(e) != 0) ? 1 : 0;
  }
  return g_deterministic != 0;
}

static int current_mode() {
  if (g_mode < 0) {
    const char* e = getenv("MINITORCH_FA_MODE");
    g_mode = (e && (!strcmp(e, "bf16") || !strcmp(e, "BF16"))) ? FA_MODE_BF16 : FA_MODE_FP32;
  }
  return g_mode;
}

// --------------------------------------------------------------------------------------------
// TMA descriptors.  cuTensorMapEncodeTiled is fetched through the runtime so the library does
// not link against libcuda.
// --------------------------------------------------------------------------------------------
// 4-D map over a (B,H,N,d) tensor with element strides (sb, sh, sn, 1); box = [box_rows][box_cols] x 1 x 1,
// 128-byte swizzle, out-of-bounds rows read as zero / are not written.
static int make_tmap(CUtensorMap* tm, const void* base, CUtensorMapDataType dt, int esize, int B, int H, int N, int d,
                     long long sb, long long sh, long long sn, int box_cols, int box_rows) {
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) return set_error(FA_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  cuuint64_t dims[4] = {(cuuint64_t)d, (cuuint64_t)N, (cuuint64_t)H, (cuuint64_t)B};
  cuuint64_t strides[3] = {(cuuint64_t)sn * esize, (cuuint64_t)sh * esize, (cuuint64_t)sb * esize};
  cuuint32_t box[4] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows, 1, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(tm, dt, 4, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return set_error(FA_ERR_CUDA, "cuTensorMapEncodeTiled failed (CUresult %d; base %p d %d N %d strides %lld %lld %lld)",
                     (int)r, base, d, N, sn, sh, sb);
  return FA_OK;
}

struct Strides {
  long long sb, sh, sn;
};
static Strides resolve_strides(const fa_attn_desc* a) {
  if (a->stride_b == 0 && a->stride_h == 0 && a->stride_n == 0)
    return {(long long)a->H * a->N * a->d, (long long)a->N * a->d, (long long)a->d};
  return {a->stride_b, a->stride_h, a->stride_n};
}

static int validate(const fa_attn_desc* a, const char* who) {
  if (!a) return set_error(FA_ERR_INVALID, "%s: null descriptor", who);
  if (a->B <= 0 || a->H <= 0 || a->N <= 0 || a->d <= 0)
    return set_error(FA_ERR_INVALID, "%s: bad shape B=%d H=%d N=%d d=%d", who, a->B, a->H, a->N, a->d);
  if (a->dtype != FA_DTYPE_F32 && a->dtype != FA_DTYPE_BF16)
    return set_error(FA_ERR_INVALID, "%s: unknown dtype %d", who, a->dtype);
  if (a->H > 65535 || a->B > 65535) return set_error(FA_ERR_UNSUPPORTED, "%s: B and H must be <= 65535", who);
  return FA_OK;
}

static AttnParams make_params(const fa_attn_desc* a) {
  Strides s = resolve_strides(a);
  AttnParams p;
  p.B = a->B, p.H = a->H, p.N = a->N, p.d = a->d;
  p.causal = a->causal ? 1 : 0;
  p.sb = s.sb, p.sh = s.sh, p.sn = s.sn;
  p.kv_len = a->kv_len;
  p.key_mask = a->key_mask;
  p.scale = 1.0f / sqrtf((float)a->d);
  return p;
}

// Head dims the tcgen05 kernels cover: every multiple of 8 up to 128.  The kernels are built for D = 64 and D = 128;
// a smaller head dim runs on the next template size with its tiles zero-padded by TMA (the box is D columns wide, the
// tensor only d: out-of-bounds columns arrive as zeros and contribute nothing to Q.K^T / P.V) and its outputs stored
// only up to d.  BASELINE config #2's model (head_dim 32) therefore runs on tensor cores too, at D = 64 cost.
static bool tc_supported(const fa_attn_desc* a) {
  if (a->dtype != FA_DTYPE_BF16) return false;
  if (a->d < 8 || a->d > 128 || (a->d % 8) != 0) return false;
  Strides s = resolve_strides(a);
  return (s.sn % 8 == 0) && (s.sh % 8 == 0) && (s.sb % 8 == 0);  // TMA strides: multiples of 16 bytes
}

// ------------------------------------------------------------------------------------- fp32 path
template <typename T>
static int fwd_simt(const fa_attn_desc* a, const void* Q, const void* K, const void* V, void* O, float* m, float* l,
                    cudaStream_t st) {
  AttnParams p = make_params(a);
  const int nqt = (p.N + 63) / 64;
  const int dcap = p.d < 256 ? p.d : 256;
  const int nch = dcap <= 64 ? 1 : (dcap <= 128 ? 2 : 4);
  const int nslice = (p.d + nch * 64 - 1) / (nch * 64);
  dim3 grid(nqt * nslice, p.H, p.B);
  auto launch = [&](auto kern) -> int {
    FA_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)f32k::kSmemFwd));
    kern<<<grid, f32k::NT, f32k::kSmemFwd, st>>>(p, (const T*)Q, (const T*)K, (const T*)V, (T*)O, m, l);
    fa::count_launch();
    FA_CUDA_CHECK(cudaGetLastError());
    return FA_OK;
  };
  if (nch == 1) return launch(f32k::fwd_kernel<T, 1>);
  if (nch == 2) return launch(f32k::fwd_kernel<T, 2>);
  return launch(f32k::fwd_kernel<T, 4>);
}

template <typename T>
static int bwd_prep(const AttnParams& p, const void* O, const void* dO, const float* m, const float* l, float* Dv,
                    float* LSE, cudaStream_t st) {
  const long long rows = (long long)p.B * p.H * p.N;
  long long blocks = (rows + 7) / 8;
  if (blocks > 148 * 32) blocks = 148 * 32;
  f32k::bwd_prep_kernel<T><<<(int)blocks, 256, 0, st>>>(p, (const T*)O, (const T*)dO, m, l, Dv, LSE);
    fa::count_launch();
  FA_CUDA_CHECK(cudaGetLastError());
  return FA_OK;
}

template <typename T>
static int bwd_simt(const fa_attn_desc* a, const void* Q, const void* K, const void* V, const void* O, const void* dO,
                    const float* m, const float* l, void* dQ, void* dK, void* dV, cudaStream_t st) {
  AttnParams p = make_params(a);
  const long long rows = (long long)p.B * p.H * p.N;
  // workspace is stream-ordered (taken from / returned to the cached default pool on `st`): calls in flight on
  // different streams never share it
  float* ws = static_cast<float*>(pool_alloc(sizeof(float) * 2 * rows, st));
  if (!ws) return set_error(FA_ERR_CUDA, "flash bwd: workspace allocation failed");
  struct Release {
    void* p;
    cudaStream_t st;
    ~Release() { cudaFreeAsync(p, st); }
  } release{ws, st};
  float *Dv = ws, *LSE = ws + rows;
  int rc = bwd_prep<T>(p, O, dO, m, l, Dv, LSE, st);
  if (rc != FA_OK) return rc;
  const int nt = (p.N + 63) / 64;
  {
    const int dcap = p.d < 128 ? p.d : 128;
    const int nch = dcap <= 64 ? 1 : 2;
    const int nslice = (p.d + nch * 64 - 1) / (nch * 64);
    dim3 grid(nt * nslice, p.H, p.B);
    auto launch = [&](auto kern) -> int {
      FA_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)f32k::kSmemBwd));
      kern<<<grid, f32k::NT, f32k::kSmemBwd, st>>>(p, (const T*)Q, (const T*)K, (const T*)V, (const T*)dO, Dv, LSE,
                                                   (T*)dK, (T*)dV);
    fa::count_launch();
      FA_CUDA_CHECK(cudaGetLastError());
      return FA_OK;
    };
    rc = (nch == 1) ? launch(f32k::bwd_dkdv_kernel<T, 1>) : launch(f32k::bwd_dkdv_kernel<T, 2>);
    if (rc != FA_OK) return rc;
  }
  {
    const int dcap = p.d < 256 ? p.d : 256;
    const int nch = dcap <= 64 ? 1 : (dcap <= 128 ? 2 : 4);
    const int nslice = (p.d + nch * 64 - 1) / (nch * 64);
    dim3 grid(nt * nslice, p.H, p.B);
    auto launch = [&](auto kern) -> int {
      FA_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)f32k::kSmemBwd));
      kern<<<grid, f32k::NT, f32k::kSmemBwd, st>>>(p, (const T*)Q, (const T*)K, (const T*)V, (const T*)dO, Dv, LSE,
                                                   (T*)dQ);
    fa::count_launch();
      FA_CUDA_CHECK(cudaGetLastError());
      return FA_OK;
    };
    if (nch == 1) return launch(f32k::bwd_dq_kernel<T, 1>);
    if (nch == 2) return launch(f32k::bwd_dq_kernel<T, 2>);
    return launch(f32k::bwd_dq_kernel<T, 4>);
  }
}

// ------------------------------------------------------------------------------ tensor-core path
template <int D, bool CAUSAL, int MASKMODE, typename OutT>
static int launch_fwd_tc(const fa_attn_desc* a, const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv,
                         const sm100::FwdParams& fp, cudaStream_t st) {
  using Cfg = sm100::FwdCfg<D>;
  // Causal problems whose whole K/V is L2-resident (the short-sequence end of BASELINE config #3) go to the persistent
  // kernel: one CTA per SM, all heads' longest query blocks first, next item's first Q.K^T under the epilogue
  // (+7..9 % at N = 512 / 1024, B8 H16; profiles/r02_fwd_persistent_ab.txt).  Everything else runs one CTA per query
  // block in head-major order: the hardware scheduler balances the unequal causal blocks dynamically and the CTAs that
  // run together share heads, so a head's K/V is read from HBM once.
  const double kv_bytes = 2.0 * a->B * a->H * (double)a->N * a->d * 2.0;
  if constexpr (CAUSAL) {
    if (kv_bytes <= 72e6 && fp.n_items > 1) {
      auto pk = sm100::fwd_persistent_kernel<D, CAUSAL, MASKMODE, OutT>;
      FA_CUDA_CHECK(cudaFuncSetAttribute(pk, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
      static int sm_count[ScratchPool::kMaxDev] = {};
      int dev = 0;
      FA_CUDA_CHECK(cudaGetDevice(&dev));
      if (dev >= 0 && dev < ScratchPool::kMaxDev && !sm_count[dev])
        FA_CUDA_CHECK(cudaDeviceGetAttribute(&sm_count[dev], cudaDevAttrMultiProcessorCount, dev));
      const int sms = (dev >= 0 && dev < ScratchPool::kMaxDev && sm_count[dev] > 0) ? sm_count[dev] : 148;
      sm100::FwdParams q = fp;
      q.qb_major = 1;
      pk<<<fp.n_items < sms ? fp.n_items : sms, Cfg::NTHREADS, Cfg::SMEM_BYTES, st>>>(tq, tk, tv, q);
      fa::count_launch();
      FA_CUDA_CHECK(cudaGetLastError());
      return FA_OK;
    }
  }
  auto kern = sm100::fwd_kernel<D, CAUSAL, MASKMODE, OutT>;
  FA_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
  dim3 grid((a->N + 255) / 256, a->H, a->B);
  kern<<<grid, Cfg::NTHREADS, Cfg::SMEM_BYTES, st>>>(tq, tk, tv, fp);
  fa::count_launch();
  FA_CUDA_CHECK(cudaGetLastError());
  return FA_OK;
}

// OutT = bf16 (the only instantiation: the legacy fp32 ABI narrows / widens around the bf16 kernels, legacy_pipeline.cuh).
template <typename OutT>
static int fwd_tc(const fa_attn_desc* a, const void* Q, const void* K, const void* V, void* O, long long o_sb,
                  long long o_sh, long long o_sn, float* m, float* l, cudaStream_t st) {
  Strides s = resolve_strides(a);
  CUtensorMap tq, tk, tv;
  int rc;
  if ((rc = make_tmap(&tq, Q, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, a->B, a->H, a->N, a->d, s.sb, s.sh, s.sn, 64, 128)))
    return rc;
  if ((rc = make_tmap(&tk, K, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, a->B, a->H, a->N, a->d, s.sb, s.sh, s.sn, 64, 128)))
    return rc;
  if ((rc = make_tmap(&tv, V, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, a->B, a->H, a->N, a->d, s.sb, s.sh, s.sn, 64, 128)))
    return rc;
  sm100::FwdParams fp;
  fp.B = a->B, fp.H = a->H, fp.N = a->N;
  fp.d = a->d;
  fp.kv_len = a->kv_len;
  fp.key_mask = a->key_mask;
  fp.O = O;
  fp.o_sb = o_sb, fp.o_sh = o_sh, fp.o_sn = o_sn;
  fp.M = m, fp.L = l;
  fp.scale = 1.0f / sqrtf((float)a->d);
  fp.scale_log2 = fp.scale * 1.4426950408889634f;
  fp.trace = g_trace;
  fp.n_qblk = (a->N + 255) / 256;
  const long long items = (long long)fp.n_qblk * a->H * a->B;
  if (items > 0x7fffffffLL) return set_error(FA_ERR_UNSUPPORTED, "flash fwd: too many work items");
  fp.n_items = (int)items;
  const int maskmode = a->key_mask ? 2 : (a->kv_len ? 1 : 0);
  const int DT = a->d <= 64 ? 64 : 128;   // template head dim
#define FA_FWD_CASE(DD, CC, MM) \
  if (DT == DD && (a->causal != 0) == CC && maskmode == MM) return launch_fwd_tc<DD, CC, MM, OutT>(a, tq, tk, tv, fp, st);
  FA_FWD_CASE(128, false, 0) FA_FWD_CASE(128, true, 0) FA_FWD_CASE(128, false, 1) FA_FWD_CASE(128, true, 1)
  FA_FWD_CASE(128, false, 2) FA_FWD_CASE(128, true, 2) FA_FWD_CASE(64, false, 0) FA_FWD_CASE(64, true, 0)
  FA_FWD_CASE(64, false, 1) FA_FWD_CASE(64, true, 1) FA_FWD_CASE(64, false, 2) FA_FWD_CASE(64, true, 2)
#undef FA_FWD_CASE
  return set_error(FA_ERR_UNSUPPORTED, "flash fwd (tensor core): head_dim %d not supported", a->d);
}

template <int D, bool CAUSAL, bool MASK2>
static int launch_bwd_tc(const fa_attn_desc* a, const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv,
                         const CUtensorMap& tdo, const CUtensorMap& tdq, const sm100::BwdParams& bp,
                         cudaStream_t st) {
  using Cfg = sm100::BwdCfg<D>;
  auto kern = sm100::bwd_kernel<D, CAUSAL, MASK2>;
  FA_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
  dim3 grid((a->N + 127) / 128, a->H, a->B);
  kern<<<grid, Cfg::NTHREADS, Cfg::SMEM_BYTES, st>>>(tq, tk, tv, tdo, tdq, bp);
  fa::count_launch();
  FA_CUDA_CHECK(cudaGetLastError());
  return FA_OK;
}

// Tensor-core backward (bf16, head_dim 64/128; no mask, kv_len and / or a generic additive key mask).
static int bwd_tc(const fa_attn_desc* a, const void* Q, const void* K, const void* V, const void* O, const void* dO,
                  const float* m, const float* l, void* dQ, void* dK, void* dV, cudaStream_t st) {
  Strides s = resolve_strides(a);
  const int Npad = ((a->N + 127) / 128) * 128;
  const size_t rows_pad = (size_t)a->B * a->H * Npad;
  const size_t nacc = (size_t)a->B * a->H * a->N * a->d;
  // stream-ordered workspace (see bwd_simt): LSE / D vectors and the fp32 dQ accumulator
  float* ws = static_cast<float*>(pool_alloc(sizeof(float) * 2 * rows_pad, st));
  float* acc = static_cast<float*>(pool_alloc(sizeof(float) * nacc, st));
  struct Release {
    void *a, *b;
    cudaStream_t st;
    ~Release() {
      if (a) cudaFreeAsync(a, st);
      if (b) cudaFreeAsync(b, st);
    }
  } release{ws, acc, st};
  if (!ws || !acc) return set_error(FA_ERR_CUDA, "flash bwd: workspace allocation failed");
  float *lse2 = ws, *dvec = ws + rows_pad;
  {
    const long long blocks = 148 * 8;
    const long long n_acc4 = (long long)(nacc / 4);
    if (a->d > 64)
      sm100::bwd_prep_tc_kernel<128><<<(int)blocks, 256, 0, st>>>(a->B, a->H, a->N, Npad, a->d, s.sb, s.sh, s.sn,
                                                                 (const __nv_bfloat16*)O, (const __nv_bfloat16*)dO, m,
                                                                 l, lse2, dvec, (float4*)acc, n_acc4);
    else
      sm100::bwd_prep_tc_kernel<64><<<(int)blocks, 256, 0, st>>>(a->B, a->H, a->N, Npad, a->d, s.sb, s.sh, s.sn,
                                                                (const __nv_bfloat16*)O, (const __nv_bfloat16*)dO, m,
                                                                l, lse2, dvec, (float4*)acc, n_acc4);
    fa::count_launch();
    FA_CUDA_CHECK(cudaGetLastError());
  }
  CUtensorMap tq, tk, tv, tdo;
  int rc;
  const void* src[4] = {Q, K, V, dO};
  CUtensorMap* tm[4] = {&tq, &tk, &tv, &tdo};
  for (int i = 0; i < 4; ++i)
    if ((rc = make_tmap(tm[i], src[i], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, a->B, a->H, a->N, a->d, s.sb, s.sh, s.sn, 64,
                        128)))
      return rc;
  // fp32 dQ accumulator (contiguous B,H,N,d): [128 rows][32 columns] boxes for the TMA add-reduction
  CUtensorMap tdq;
  if ((rc = make_tmap(&tdq, acc, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, a->B, a->H, a->N, a->d,
                      (long long)a->H * a->N * a->d, (long long)a->N * a->d, a->d, 32, 128)))
    return rc;
  sm100::BwdParams bp;
  bp.B = a->B, bp.H = a->H, bp.N = a->N, bp.Npad = Npad;
  bp.d = a->d;
  bp.kv_len = a->kv_len;
  bp.key_mask = a->key_mask;
  bp.lse2 = lse2, bp.dvec = dvec, bp.dq_acc = acc;
  bp.dK = dK, bp.dV = dV;
  bp.sb = s.sb, bp.sh = s.sh, bp.sn = s.sn;
  bp.scale = 1.0f / sqrtf((float)a->d);
  bp.scale_log2 = bp.scale * 1.4426950408889634f;
  bp.trace = g_trace;
  rc = FA_ERR_UNSUPPORTED;
  const int DT = a->d <= 64 ? 64 : 128;   // template head dim (see tc_supported)
#define FA_BWD_CASE(DD, CC, MM)                                                       \
  if (DT == DD && (a->causal != 0) == CC && (a->key_mask != nullptr) == MM)           \
    rc = launch_bwd_tc<DD, CC, MM>(a, tq, tk, tv, tdo, tdq, bp, st);
  FA_BWD_CASE(128, false, false) FA_BWD_CASE(128, true, false) FA_BWD_CASE(128, false, true) FA_BWD_CASE(128, true, true)
  FA_BWD_CASE(64, false, false) FA_BWD_CASE(64, true, false) FA_BWD_CASE(64, false, true) FA_BWD_CASE(64, true, true)
#undef FA_BWD_CASE
  if (rc) return rc;
  {
    const long long total8 = (long long)nacc / 8;
    long long blocks = (total8 + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    sm100::bwd_convert_dq_kernel<<<(int)blocks, 256, 0, st>>>(a->H, a->N, a->d, s.sb, s.sh, s.sn, bp.scale, acc,
                                                              (__nv_bfloat16*)dQ, total8);
    fa::count_launch();
    FA_CUDA_CHECK(cudaGetLastError());
  }
  return FA_OK;
}

__global__ void cast_f32_bf16_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, size_t n) {
  size_t i = (static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x) * 8;
  const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x * 8;
  for (; i + 8 <= n; i += stride) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(src + i));
    const float4 b = __ldg(reinterpret_cast<const float4*>(src + i + 4));
    uint4 o;
    o.x = pack_bf16x2(a.x, a.y), o.y = pack_bf16x2(a.z, a.w), o.z = pack_bf16x2(b.x, b.y), o.w = pack_bf16x2(b.z, b.w);
    *reinterpret_cast<uint4*>(dst + i) = o;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0)
    for (size_t t = n & ~size_t(7); t < n; ++t) dst[t] = __float2bfloat16_rn(src[t]);
}
__global__ void cast_bf16_f32_kernel(const __nv_bfloat16* __restrict__ src, float* __restrict__ dst, size_t n) {
  size_t i = (static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x) * 8;
  const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x * 8;
  for (; i + 8 <= n; i += stride) {
    const uint4 v = __ldg(reinterpret_cast<const uint4*>(src + i));
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
    float o[8];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      o[2 * k] = __uint_as_float(w[k] << 16);
      o[2 * k + 1] = __uint_as_float(w[k] & 0xffff0000u);
    }
    *reinterpret_cast<float4*>(dst + i) = make_float4(o[0], o[1], o[2], o[3]);
    *reinterpret_cast<float4*>(dst + i + 4) = make_float4(o[4], o[5], o[6], o[7]);
  }
  if (blockIdx.x == 0 && threadIdx.x == 0)
    for (size_t t = n & ~size_t(7); t < n; ++t) dst[t] = __bfloat162float(src[t]);
}
static int cast_grid(size_t n) {
  size_t b = (n / 8 + 255) / 256;
  if (b < 1) b = 1;
  if (b > 148 * 16) b = 148 * 16;
  return (int)b;
}

}  // namespace fa

using namespace fa;

extern "C" {

void fa_set_mode(int mode) {
  current_mode();
  g_mode = (mode == FA_MODE_BF16) ? FA_MODE_BF16 : FA_MODE_FP32;
}
int fa_get_mode(void) { return current_mode(); }
void fa_set_deterministic(int on) { g_deterministic = on ? 1 : 0; }
#ifdef FA_TRACE
void fa_debug_set_trace(long long* dev_buf) { g_trace = dev_buf; }
#endif

int fa_cast_f32_to_bf16_dev(const float* src, void* dst, size_t n, fa_stream_t stream) {
  clear_error();
  if (n == 0) return FA_OK;
  cast_f32_bf16_kernel<<<cast_grid(n), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      src, static_cast<__nv_bfloat16*>(dst), n);
    fa::count_launch();
  FA_CUDA_CHECK(cudaGetLastError());
  return FA_OK;
}
int fa_cast_bf16_to_f32_dev(const void* src, float* dst, size_t n, fa_stream_t stream) {
  clear_error();
  if (n == 0) return FA_OK;
  cast_bf16_f32_kernel<<<cast_grid(n), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(src), dst, n);
    fa::count_launch();
  FA_CUDA_CHECK(cudaGetLastError());
  return FA_OK;
}

double fa_attn_flops(int B, int H, int N, int d, int causal, const int* kv_len_host, int backward) {
  // fwd 4*B*H*Nq*Nk*d, bwd 10*..., causal halves, padding replaces Nk by kv_len[b].
  double pairs = 0.0;
  for (int b = 0; b < B; ++b) {
    const double nk = kv_len_host ? (double)(kv_len_host[b] < N ? kv_len_host[b] : N) : (double)N;
    // causal: query i sees min(i+1, nk) keys; counted with the usual "x 1/2" convention
    pairs += causal ? (0.5 * nk * nk + ((double)N - nk) * nk) : (double)N * nk;
  }
  return (backward ? 10.0 : 4.0) * (double)H * (double)d * pairs;
}

int fa_flash_fwd_dev(const fa_attn_desc* a, const void* Q, const void* K, const void* V, void* O, float* m, float* l,
                     fa_stream_t stream) {
  clear_error();
  int rc = validate(a, "fa_flash_fwd_dev");
  if (rc) return rc;
  if (!Q || !K || !V || !O || !m || !l) return set_error(FA_ERR_INVALID, "fa_flash_fwd_dev: null tensor pointer");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (a->dtype == FA_DTYPE_F32) return fwd_simt<float>(a, Q, K, V, O, m, l, st);
  if (tc_supported(a)) {
    Strides s = resolve_strides(a);
    return fwd_tc<__nv_bfloat16>(a, Q, K, V, O, s.sb, s.sh, s.sn, m, l, st);
  }
  // bf16 tensors with a head dim the tcgen05 kernels do not tile: CUDA-core kernel, fp32 math
  return fwd_simt<__nv_bfloat16>(a, Q, K, V, O, m, l, st);
}

int fa_flash_bwd_dev(const fa_attn_desc* a, const void* Q, const void* K, const void* V, const void* O, const void* dO,
                     const float* m, const float* l, void* dQ, void* dK, void* dV, fa_stream_t stream) {
  clear_error();
  int rc = validate(a, "fa_flash_bwd_dev");
  if (rc) return rc;
  if (!Q || !K || !V || !O || !dO || !m || !l || !dQ || !dK || !dV)
    return set_error(FA_ERR_INVALID, "fa_flash_bwd_dev: null tensor pointer");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (a->dtype == FA_DTYPE_F32) return bwd_simt<float>(a, Q, K, V, O, dO, m, l, dQ, dK, dV, st);
  if (tc_supported(a) && !deterministic()) {
    int r2 = bwd_tc(a, Q, K, V, O, dO, m, l, dQ, dK, dV, st);
    if (r2 != FA_ERR_UNSUPPORTED) return r2;
    clear_error();
  }
  return bwd_simt<__nv_bfloat16>(a, Q, K, V, O, dO, m, l, dQ, dK, dV, st);
}

// ---------------------------------------------------------------------------------------------
// Decode shapes: one query token per (batch, head) against a KV cache (flash_decode.cuh).
// ---------------------------------------------------------------------------------------------
}  // extern "C"
template <typename T>
static int launch_decode(const fa_decode_desc* a, const decode::Params& p, const void* q, const void* kc, const void* vc,
                         void* out, cudaStream_t st) {
  constexpr int VN = decode::Vec<T>::N;
  const int nvec = a->d / VN;
  dim3 grid(p.nsplit, a->H, a->B);
  const size_t smem = sizeof(float) * decode::kWarps * a->d;
#define FA_DEC_CASE(LPK, NV)                                                                                        \
  decode::partial_kernel<T, LPK, NV><<<grid, decode::kWarps * 32, smem, st>>>(p, (const T*)q, (const T*)kc,        \
                                                                               (const T*)vc, (T*)out)
  if (nvec <= 1) FA_DEC_CASE(1, 1);
  else if (nvec <= 2) FA_DEC_CASE(2, 1);
  else if (nvec <= 4) FA_DEC_CASE(4, 1);
  else if (nvec <= 8) FA_DEC_CASE(8, 1);
  else if (nvec <= 16) FA_DEC_CASE(16, 1);
  else if (nvec <= 32) FA_DEC_CASE(32, 1);
  else FA_DEC_CASE(32, 2);
#undef FA_DEC_CASE
  fa::count_launch();
  FA_CUDA_CHECK(cudaGetLastError());
  if (p.nsplit > 1) {
    decode::combine_kernel<T><<<dim3(a->H, a->B), 128, 0, st>>>(p, (T*)out);
    fa::count_launch();
    FA_CUDA_CHECK(cudaGetLastError());
  }
  return FA_OK;
}
extern "C" {

int fa_flash_decode_dev(const fa_decode_desc* a, const void* q, const void* k_cache, const void* v_cache, void* out,
                        float* lse, fa_stream_t stream) {
  clear_error();
  if (!a || !q || !k_cache || !v_cache || !out) return set_error(FA_ERR_INVALID, "fa_flash_decode_dev: null argument");
  if (a->B <= 0 || a->H <= 0 || a->d <= 0 || a->L < 0 || a->B > 65535 || a->H > 65535)
    return set_error(FA_ERR_INVALID, "fa_flash_decode_dev: bad shape B=%d H=%d d=%d L=%d", a->B, a->H, a->d, a->L);
  if (a->dtype != FA_DTYPE_F32 && a->dtype != FA_DTYPE_BF16)
    return set_error(FA_ERR_INVALID, "fa_flash_decode_dev: unknown dtype %d", a->dtype);
  const int VN = a->dtype == FA_DTYPE_BF16 ? 8 : 4;
  const int cap = a->L_cap > 0 ? a->L_cap : a->L;
  decode::Params p{};
  p.B = a->B, p.H = a->H, p.d = a->d, p.L = a->L, p.kv_len = a->kv_len;
  p.q_sb = a->q_stride_b ? a->q_stride_b : (long long)a->H * a->d;
  p.q_sh = a->q_stride_h ? a->q_stride_h : a->d;
  p.o_sb = a->o_stride_b ? a->o_stride_b : (long long)a->H * a->d;
  p.o_sh = a->o_stride_h ? a->o_stride_h : a->d;
  const bool cdef = !a->cache_stride_b && !a->cache_stride_h && !a->cache_stride_n;
  p.c_sb = cdef ? (long long)a->H * cap * a->d : a->cache_stride_b;
  p.c_sh = cdef ? (long long)cap * a->d : a->cache_stride_h;
  p.c_sn = cdef ? a->d : a->cache_stride_n;
  if (a->d % VN || a->d / VN > 64 || p.q_sb % VN || p.q_sh % VN || p.c_sb % VN || p.c_sh % VN || p.c_sn % VN)
    return set_error(FA_ERR_UNSUPPORTED,
                     "fa_flash_decode_dev: head_dim %d / strides must be multiples of %d elements (16 bytes) and head_dim "
                     "<= %d", a->d, VN, 64 * VN);
  p.scale_log2 = 1.4426950408889634f / sqrtf((float)a->d);
  p.lse = lse;
  // Split the cache so that the CTAs fill ONE co-resident wave and no more (a second, partial wave costs a whole
  // tail): 148 SMs x the CTAs of 4 warps that fit one (80 registers with bf16 caches -> 6, 64 with fp32 -> 8), at
  // least 128 keys per split.  MINITORCH_FA_DECODE_SPLITS forces a count (experiments).
  static const int forced = [] {
    const char* e = getenv("MINITORCH_FA_DECODE_SPLITS");
    return e ? atoi(e) : 0;
  }();
  const int Lmax = a->kv_len ? cap : a->L;
  const int slots = 148 * (a->dtype == FA_DTYPE_BF16 ? 6 : 8);
  int nsplit = forced > 0 ? forced : (int)(slots / ((long long)a->B * a->H));
  const int max_split = (Lmax + 127) / 128;
  nsplit = nsplit > max_split ? max_split : nsplit;
  nsplit = nsplit < 1 ? 1 : (nsplit > 64 ? 64 : nsplit);
  p.nsplit = nsplit;
  p.chunk = (((Lmax + nsplit - 1) / nsplit) + 15) & ~15;
  if (p.chunk < 16) p.chunk = 16;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  float* ws = nullptr;
  if (nsplit > 1) {
    const size_t units = (size_t)a->B * a->H * nsplit;
    ws = static_cast<float*>(pool_alloc(sizeof(float) * units * (2 + a->d), st));
    if (!ws) return set_error(FA_ERR_CUDA, "fa_flash_decode_dev: workspace allocation failed");
    p.ws_ml = ws;
    p.ws_o = ws + units * 2;
  }
  const int rc = (a->dtype == FA_DTYPE_BF16) ? launch_decode<__nv_bfloat16>(a, p, q, k_cache, v_cache, out, st)
                                             : launch_decode<float>(a, p, q, k_cache, v_cache, out, st);
  if (ws) cudaFreeAsync(ws, st);
  return rc;
}

}  // extern "C" (device-pointer API above; the legacy host-pointer ABI follows)

namespace fa {
static bool tc_head_dim(int d) { return d >= 8 && d <= 128 && d % 8 == 0; }
static int fwd_tc_bf16(const fa_attn_desc* a, const void* Q, const void* K, const void* V, void* O, float* m, float* l,
                       cudaStream_t st) {
  Strides s = resolve_strides(a);
  return fwd_tc<__nv_bfloat16>(a, Q, K, V, O, s.sb, s.sh, s.sn, m, l, st);
}
}  // namespace fa
#include "legacy_pipeline.cuh"

extern "C" {

// ---------------------------------------------------------------------------------------------
// Legacy host-pointer ABI.  fp32 host buffers in, fp32 host buffers out (legacy_pipeline.cuh).
// ---------------------------------------------------------------------------------------------
void fa_set_legacy_chunk_bytes(size_t bytes) {
  g_chunk_bytes = bytes ? bytes : ((size_t)16 << 20);
  g_chunk_explicit = bytes != 0;
}
void fa_set_keep_forward_mb(long long mb) { g_keep_budget = (mb < 0 ? 0 : mb) << 20; }
void fa_forward_cache_stats(unsigned long long* hits, unsigned long long* misses) {
  if (hits) *hits = g_fwd_hits;
  if (misses) *misses = g_fwd_misses;
}
unsigned long long fa_staging_fallbacks(void) { return g_staging_fallbacks; }
void fa_set_transfer_policy(double host_cost, long long min_tensor_bytes) {
  g_hybrid_cost = host_cost;
  g_hybrid_min_bytes = min_tensor_bytes < 0 ? ((size_t)32 << 20) : (size_t)min_tensor_bytes;
}
// The split plan_hybrid() would choose for a call with n_in uploaded and n_out downloaded large bf16-mode tensors
// (pageable[i] != 0: that tensor is pageable and must go through the staging threads) at the given host cost:
// writes the quarters of each tensor that take the staged route (0..4).  Pure host logic, no CUDA call.
int fa_plan_transfer_preview(int n_in, const int* in_pageable, int n_out, const int* out_pageable, double host_cost,
                             int* in_staged_q, int* out_staged_q) {
  if (n_in < 0 || n_out < 0 || n_in > 16 || n_out > 16 || (n_in && (!in_pageable || !in_staged_q)) ||
      (n_out && (!out_pageable || !out_staged_q)))
    return set_error(FA_ERR_INVALID, "fa_plan_transfer_preview: bad arguments");
  std::vector<Xfer> ins(n_in), outs(n_out);
  const size_t unit = (size_t)64 << 20;     // one unit per tensor, far above the split threshold
  auto fill = [&](std::vector<Xfer>& xs, const int* pageable) {
    for (size_t i = 0; i < xs.size(); ++i) {
      xs[i].unit = unit, xs[i].wire = WIRE_BF16;
      xs[i].direct = !pageable[i];
      xs[i].staged_q = pageable[i] ? 4 : 0;
    }
  };
  fill(ins, in_pageable), fill(outs, out_pageable);
  plan_hybrid(ins, outs, 1, host_cost);
  for (int i = 0; i < n_in; ++i) in_staged_q[i] = ins[i].staged_q;
  for (int i = 0; i < n_out; ++i) out_staged_q[i] = outs[i].staged_q;
  return FA_OK;
}
void fa_wire_bytes(unsigned long long* h2d, unsigned long long* d2h) {
  if (h2d) *h2d = g_wire_h2d.load();
  if (d2h) *d2h = g_wire_d2h.load();
}
// Gives back what the legacy entry points hold between calls on the current device: the pinned staging rings and the
// device copies of forward calls that are still waiting for their backward.
int fa_release_staging(void) {
  clear_error();
  int dev = 0;
  FA_CUDA_CHECK(cudaGetDevice(&dev));
  if (dev < 0 || dev >= ScratchPool::kMaxDev) return FA_OK;
  DevPipe& P = g_pipes[dev];
  if (P.ready) {
    cudaError_t e = P.drain();
    if (e != cudaSuccess) return set_error(FA_ERR_CUDA, "fa_release_staging: %s", cudaGetErrorString(e));
  }
  P.rin.release();
  P.rout.release();
  drop_cached_forwards(dev, nullptr);
  FA_CUDA_CHECK(cudaStreamSynchronize(nullptr));
  return FA_OK;
}
void launch_flashattention_forward(float* Q, float* K, float* V, float* O, float* l, float* m, int B, int nh, int N,
                                   int d) {
  legacy_forward(Q, K, V, O, l, m, nullptr, 0, B, nh, N, d);
}
void launch_flashattention_forward_causal(float* Q, float* K, float* V, float* O, float* l, float* m, int B, int nh,
                                          int N, int d) {
  legacy_forward(Q, K, V, O, l, m, nullptr, 1, B, nh, N, d);
}
void launch_flashattention_forward_masked(float* Q, float* K, float* V, float* O, float* l, float* m,
                                          const float* key_mask, int causal, int B, int nh, int N, int d) {
  legacy_forward(Q, K, V, O, l, m, key_mask, causal ? 1 : 0, B, nh, N, d);
}
void launch_flashattention_backward(float* Q, float* K, float* V, float* O, float* dQ, float* dK, float* dV, float* dO,
                                    float* l, float* m, int B, int nh, int N, int d) {
  legacy_backward(Q, K, V, O, dQ, dK, dV, dO, l, m, nullptr, 0, B, nh, N, d);
}
void launch_flashattention_backward_causal(float* Q, float* K, float* V, float* O, float* dQ, float* dK, float* dV,
                                           float* dO, float* l, float* m, int B, int nh, int N, int d) {
  legacy_backward(Q, K, V, O, dQ, dK, dV, dO, l, m, nullptr, 1, B, nh, N, d);
}
void launch_flashattention_backward_masked(float* Q, float* K, float* V, float* O, float* dQ, float* dK, float* dV,
                                           float* dO, float* l, float* m, const float* key_mask, int causal, int B,
                                           int nh, int N, int d) {
  legacy_backward(Q, K, V, O, dQ, dK, dV, dO, l, m, key_mask, causal ? 1 : 0, B, nh, N, d);
}

}  // extern "C"
