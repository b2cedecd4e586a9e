// Host-side plumbing shared by the four shared libraries (included once per .so):
// status/error reporting, a grow-only device scratch pool for the legacy host-pointer
// entry points (the reference does cudaMalloc/cudaFree on every call,
// src/flashattention_kernel.cu:280-324), and the fa_* utility exports declared in
// include/flashattn_b200.h.
#pragma once
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cuda_runtime.h>

#include "../../include/flashattn_b200.h"

namespace fa {

static thread_local int g_status = FA_OK;
static thread_local char g_errmsg[512] = "";

inline int set_error(int code, const char* fmt, ...) {
  g_status = code;
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_errmsg, sizeof(g_errmsg), fmt, ap);
  va_end(ap);
  return code;
}
inline void clear_error() {
  g_status = FA_OK;
  g_errmsg[0] = 0;
}

#define FA_CUDA_CHECK(expr)                                                                      \
  do {                                                                                           \
    cudaError_t _e = (expr);                                                                     \
    if (_e != cudaSuccess)                                                                       \
      return fa::set_error(FA_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e),  \
                           __FILE__, __LINE__);                                                  \
  } while (0)

// Grow-only scratch slots, one set per device.  Single-threaded callers (ctypes from Python).
struct ScratchPool {
  static constexpr int kSlots = 32;
  static constexpr int kMaxDev = 16;
  void* ptr[kMaxDev][kSlots] = {};
  size_t cap[kMaxDev][kSlots] = {};
  void* get(int slot, size_t bytes) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= kMaxDev || slot < 0 || slot >= kSlots) return nullptr;
    if (bytes == 0) bytes = 16;
    if (cap[dev][slot] >= bytes) return ptr[dev][slot];
    if (ptr[dev][slot]) cudaFree(ptr[dev][slot]);
    ptr[dev][slot] = nullptr;
    cap[dev][slot] = 0;
    size_t want = bytes + bytes / 8 + 256;
    if (cudaMalloc(&ptr[dev][slot], want) != cudaSuccess) {
      cudaGetLastError();
      return nullptr;
    }
    cap[dev][slot] = want;
    return ptr[dev][slot];
  }
};
static ScratchPool g_pool;

// Stream-ordered allocation from the device's default memory pool.  The release threshold is raised once per device so
// that freed blocks stay cached in the pool: a workspace taken and returned on every call costs no driver round trip,
// and two calls in flight on different streams can never share a buffer (which the grow-only slots above cannot promise).
inline void tune_default_pool() {
  static bool tuned[ScratchPool::kMaxDev] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= ScratchPool::kMaxDev || tuned[dev]) return;
  cudaMemPool_t pool;
  if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
    unsigned long long keep = ~0ull;
    cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
  }
  tuned[dev] = true;
}
inline void* pool_alloc(size_t bytes, cudaStream_t st) {
  tune_default_pool();
  void* p = nullptr;
  if (cudaMallocAsync(&p, bytes ? bytes : 16, st) != cudaSuccess) {
    cudaGetLastError();
    return nullptr;
  }
  return p;
}

// second half of fa_flush_l2: stream 256 MB of clean data through L2 (see there)
__global__ void flush_read_kernel(const uint4* __restrict__ p, size_t n, unsigned* __restrict__ sink) {
  unsigned acc = 0;
  for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const uint4 v = __ldcg(p + i);
    acc ^= v.x ^ v.y ^ v.z ^ v.w;
  }
  if (acc == 0x9e3779b9u) *sink = acc;   // never true for the zero-filled buffer: keeps the loads alive
}

// Number of kernels this library has launched (bench.py reports it as gpu_launches).
static unsigned long long g_launches = 0;
inline void count_launch(int n = 1) { g_launches += static_cast<unsigned long long>(n); }

}  // namespace fa

extern "C" {
int fa_last_status(void) { return fa::g_status; }
unsigned long long fa_launch_count(void) { return fa::g_launches; }
const char* fa_last_error(void) { return fa::g_errmsg; }
int fa_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}
int fa_set_device(int dev) {
  fa::clear_error();
  FA_CUDA_CHECK(cudaSetDevice(dev));
  return FA_OK;
}
void* fa_malloc(size_t bytes) {
  fa::clear_error();
  void* p = nullptr;
  cudaError_t e = cudaMalloc(&p, bytes ? bytes : 16);
  if (e != cudaSuccess) {
    fa::set_error(FA_ERR_CUDA, "cudaMalloc(%zu) failed: %s", bytes, cudaGetErrorString(e));
    return nullptr;
  }
  return p;
}
void* fa_malloc_host(size_t bytes) {
  fa::clear_error();
  void* p = nullptr;
  cudaError_t e = cudaMallocHost(&p, bytes ? bytes : 16);
  if (e != cudaSuccess) {
    fa::set_error(FA_ERR_CUDA, "cudaMallocHost(%zu) failed: %s", bytes, cudaGetErrorString(e));
    return nullptr;
  }
  return p;
}
// Stream-ordered allocation from the device's default memory pool (release threshold raised so freed blocks stay
// cached): what a device-resident tensor library allocates its per-op outputs from.
void* fa_malloc_async(size_t bytes, fa_stream_t stream) {
  fa::clear_error();
  void* p = fa::pool_alloc(bytes, reinterpret_cast<cudaStream_t>(stream));
  if (!p) fa::set_error(FA_ERR_CUDA, "cudaMallocAsync(%zu) failed", bytes);
  return p;
}
int fa_free_async(void* p, fa_stream_t stream) {
  fa::clear_error();
  FA_CUDA_CHECK(cudaFreeAsync(p, reinterpret_cast<cudaStream_t>(stream)));
  return FA_OK;
}
int fa_memset_async(void* p, int byte, size_t bytes, fa_stream_t stream) {
  fa::clear_error();
  FA_CUDA_CHECK(cudaMemsetAsync(p, byte, bytes, reinterpret_cast<cudaStream_t>(stream)));
  return FA_OK;
}
int fa_free(void* p) {
  fa::clear_error();
  FA_CUDA_CHECK(cudaFree(p));
  return FA_OK;
}
int fa_free_host(void* p) {
  fa::clear_error();
  FA_CUDA_CHECK(cudaFreeHost(p));
  return FA_OK;
}
int fa_memset(void* p, int byte, size_t bytes) {
  fa::clear_error();
  FA_CUDA_CHECK(cudaMemset(p, byte, bytes));
  return FA_OK;
}
int fa_h2d(void* dst, const void* src, size_t bytes) {
  fa::clear_error();
  FA_CUDA_CHECK(cudaMemcpy(dst, src, bytes, cudaMemcpyHostToDevice));
  return FA_OK;
}
int fa_d2h(void* dst, const void* src, size_t bytes) {
  fa::clear_error();
  FA_CUDA_CHECK(cudaMemcpy(dst, src, bytes, cudaMemcpyDeviceToHost));
  return FA_OK;
}
int fa_sync(void) {
  fa::clear_error();
  FA_CUDA_CHECK(cudaDeviceSynchronize());
  return FA_OK;
}
void* fa_event_create(void) {
  cudaEvent_t e = nullptr;
  if (cudaEventCreate(&e) != cudaSuccess) {
    cudaGetLastError();
    return nullptr;
  }
  return e;
}
int fa_event_record(void* ev, fa_stream_t s) {
  fa::clear_error();
  FA_CUDA_CHECK(cudaEventRecord(static_cast<cudaEvent_t>(ev), reinterpret_cast<cudaStream_t>(s)));
  return FA_OK;
}
float fa_event_elapsed_ms(void* a, void* b) {
  fa::clear_error();
  float ms = -1.f;
  if (cudaEventSynchronize(static_cast<cudaEvent_t>(b)) != cudaSuccess ||
      cudaEventElapsedTime(&ms, static_cast<cudaEvent_t>(a), static_cast<cudaEvent_t>(b)) != cudaSuccess) {
    fa::set_error(FA_ERR_CUDA, "event timing failed: %s", cudaGetErrorString(cudaGetLastError()));
    return -1.f;
  }
  return ms;
}
int fa_event_destroy(void* ev) {
  fa::clear_error();
  FA_CUDA_CHECK(cudaEventDestroy(static_cast<cudaEvent_t>(ev)));
  return FA_OK;
}
// Overwrite a buffer larger than the 126 MB L2, then stream a second, clean buffer of the same size through it: after
// the write pass alone L2 is full of DIRTY lines whose write-back would be charged to the next (timed) kernel --
// +126 MB of DRAM traffic, half the traffic of a 70 us kernel.  Both passes run on the default stream.
int fa_flush_l2(void) {
  fa::clear_error();
  const size_t bytes = 256ull << 20;  // > 126 MB L2
  char* p = static_cast<char*>(fa::g_pool.get(fa::ScratchPool::kSlots - 1, 2 * bytes + 256));
  if (!p) return fa::set_error(FA_ERR_CUDA, "flush buffer allocation failed");
  static bool zeroed[fa::ScratchPool::kMaxDev] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev >= 0 && dev < fa::ScratchPool::kMaxDev && !zeroed[dev]) {
    FA_CUDA_CHECK(cudaMemsetAsync(p + bytes, 0, bytes + 256, 0));
    zeroed[dev] = true;
  }
  FA_CUDA_CHECK(cudaMemsetAsync(p, 1, bytes, 0));
  fa::flush_read_kernel<<<148 * 8, 256>>>(reinterpret_cast<const uint4*>(p + bytes), bytes / 16,
                                          reinterpret_cast<unsigned*>(p + 2 * bytes));
  FA_CUDA_CHECK(cudaGetLastError());
  return FA_OK;
}
}  // extern "C"
