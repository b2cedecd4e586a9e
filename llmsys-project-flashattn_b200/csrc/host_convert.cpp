// Host-side fp32 <-> bf16 conversion for the legacy (host-pointer) flash-attention ABI: in FA_MODE_BF16 the wire format
// between host and device is bf16, and the narrowing / widening is done by the staging threads while they copy a chunk
// between caller memory and the pinned ring (legacy_pipeline.cuh).  That copy is the end-to-end bottleneck once several
// ranks share a host (8 ranks x 6.4 GB touched per step on 32 cores), so it is vectorised: AVX2 when the CPU has it
// (16 elements per iteration, chosen at run time), scalar otherwise.  Plain C++ compiled by the host compiler and linked
// into flashattention_kernel.so (compile_cuda.sh); bit-identical to cvt.rn.bf16.f32 (round to nearest even, NaN stays NaN).
#include <cstddef>
#include <cstdint>
#include <cstdlib>
#include <immintrin.h>

namespace {

bool stream_stores();

inline uint16_t narrow_one(uint32_t u) {
  if ((u & 0x7fffffffu) > 0x7f800000u) return static_cast<uint16_t>((u >> 16) | 0x40u);
  return static_cast<uint16_t>((u + 0x7fffu + ((u >> 16) & 1u)) >> 16);
}

void narrow_scalar(uint16_t* dst, const uint32_t* src, size_t n) {
  for (size_t i = 0; i < n; ++i) dst[i] = narrow_one(src[i]);
}
void widen_scalar(uint32_t* dst, const uint16_t* src, size_t n) {
  for (size_t i = 0; i < n; ++i) dst[i] = static_cast<uint32_t>(src[i]) << 16;
}

__attribute__((target("avx2"))) inline __m256i narrow8(__m256i u) {
  const __m256i abs = _mm256_and_si256(u, _mm256_set1_epi32(0x7fffffff));
  const __m256i is_nan = _mm256_cmpgt_epi32(abs, _mm256_set1_epi32(0x7f800000));
  const __m256i lsb = _mm256_and_si256(_mm256_srli_epi32(u, 16), _mm256_set1_epi32(1));
  const __m256i rounded = _mm256_srli_epi32(_mm256_add_epi32(u, _mm256_add_epi32(lsb, _mm256_set1_epi32(0x7fff))), 16);
  const __m256i quiet = _mm256_or_si256(_mm256_srli_epi32(u, 16), _mm256_set1_epi32(0x40));
  return _mm256_blendv_epi8(rounded, quiet, is_nan);   // 8 x (0x0000hhhh)
}

// The destinations of both conversions are written once and not read again by this core (the staging slot is read by
// the DMA engine, the caller's result array by the caller much later), so 32-byte aligned destinations get streaming
// stores: no read-for-ownership of the destination lines, a third less memory traffic for the narrowing pass.
__attribute__((target("avx2"))) void narrow_avx2(uint16_t* dst, const uint32_t* src, size_t n) {
  size_t i = 0;
  const bool stream = stream_stores() && (reinterpret_cast<uintptr_t>(dst) & 31) == 0;
  for (; i + 16 <= n; i += 16) {
    const __m256i a = narrow8(_mm256_loadu_si256(reinterpret_cast<const __m256i*>(src + i)));
    const __m256i b = narrow8(_mm256_loadu_si256(reinterpret_cast<const __m256i*>(src + i + 8)));
    // packus works per 128-bit lane: [a0-3 b0-3 | a4-7 b4-7] -> permute the 64-bit quarters back into order
    const __m256i p = _mm256_permute4x64_epi64(_mm256_packus_epi32(a, b), 0xD8);
    if (stream) _mm256_stream_si256(reinterpret_cast<__m256i*>(dst + i), p);
    else _mm256_storeu_si256(reinterpret_cast<__m256i*>(dst + i), p);
  }
  if (stream) _mm_sfence();   // streaming stores are ordered before whatever publishes the buffer (the DMA doorbell)
  narrow_scalar(dst + i, src + i, n - i);
}

__attribute__((target("avx2"))) void widen_avx2(uint32_t* dst, const uint16_t* src, size_t n) {
  size_t i = 0;
  const bool stream = stream_stores() && (reinterpret_cast<uintptr_t>(dst) & 31) == 0;
  for (; i + 16 <= n; i += 16) {
    const __m256i h = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(src + i));
    const __m256i lo = _mm256_slli_epi32(_mm256_cvtepu16_epi32(_mm256_castsi256_si128(h)), 16);
    const __m256i hi = _mm256_slli_epi32(_mm256_cvtepu16_epi32(_mm256_extracti128_si256(h, 1)), 16);
    if (stream) {
      _mm256_stream_si256(reinterpret_cast<__m256i*>(dst + i), lo);
      _mm256_stream_si256(reinterpret_cast<__m256i*>(dst + i + 8), hi);
    } else {
      _mm256_storeu_si256(reinterpret_cast<__m256i*>(dst + i), lo);
      _mm256_storeu_si256(reinterpret_cast<__m256i*>(dst + i + 8), hi);
    }
  }
  if (stream) _mm_sfence();
  widen_scalar(dst + i, src + i, n - i);
}

// streaming (non-temporal) stores for 32-byte aligned destinations: env MINITORCH_FA_STREAM_STORES=0 turns them off
bool stream_stores() {
  static const bool v = [] {
    const char* e = getenv("MINITORCH_FA_STREAM_STORES");
    return !(e && e[0] == '0');
  }();
  return v;
}

bool have_avx2() {
  static const bool v = __builtin_cpu_supports("avx2");
  return v;
}

}  // namespace

extern "C" {
// fp32 -> bf16 bit patterns (round to nearest even), n elements
void fa_host_narrow_f32_bf16(uint16_t* dst, const float* src, size_t n) {
  if (have_avx2()) narrow_avx2(dst, reinterpret_cast<const uint32_t*>(src), n);
  else narrow_scalar(dst, reinterpret_cast<const uint32_t*>(src), n);
}
// bf16 bit patterns -> fp32, n elements
void fa_host_widen_bf16_f32(float* dst, const uint16_t* src, size_t n) {
  if (have_avx2()) widen_avx2(reinterpret_cast<uint32_t*>(dst), src, n);
  else widen_scalar(reinterpret_cast<uint32_t*>(dst), src, n);
}
int fa_host_convert_isa(void) { return have_avx2() ? 2 : 0; }
}
