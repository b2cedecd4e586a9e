// Micro-benchmark: TMEM -> register (tcgen05.ld) and register -> TMEM (tcgen05.st) throughput,
// and MUFU.EX2 throughput, on one SM.  Bring-up tool (not part of the product).
#include <cstdio>
#include "../../llmsys-project-flashattn_b200/csrc/ptx.cuh"
using namespace fa;

template <int MODE>
__global__ void __launch_bounds__(256, 1) k(long long* out, int iters) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) tmem_alloc<512>(&slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t base = slot + (static_cast<uint32_t>((warp & 3) * 32) << 16);
  uint32_t acc = 0;
  uint32_t u[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) u[i] = threadIdx.x + i;
  float f = threadIdx.x * 0.001f;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    if (MODE == 0) {  // load 128 columns (4 x32), one wait
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        uint32_t v[32];
        tmem_ld32(base + 32 * c, v);
        tmem_wait_ld();
#pragma unroll
        for (int i = 0; i < 32; ++i) acc ^= v[i];
      }
    } else if (MODE == 1) {  // store 64 columns (2 x32)
#pragma unroll
      for (int c = 0; c < 2; ++c) tmem_st32(base + 32 * c, u);
      tmem_wait_st();
    } else {  // 128 ex2
#pragma unroll
      for (int i = 0; i < 128; ++i) f = ex2_approx(f) * 0.5f;
    }
  }
  long long t1 = clock64();
  if (threadIdx.x % 32 == 0) out[warp] = t1 - t0;
  if (acc == 0x12345678u || f == 123.f) out[20] = acc;
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(slot);
}

int main() {
  long long* d;
  cudaMalloc(&d, 64 * 8);
  long long h[64];
  const int iters = 2000;
  for (int nthreads : {128, 256}) {
    for (int mode = 0; mode < 3; ++mode) {
      if (mode == 0) k<0><<<1, nthreads>>>(d, iters);
      if (mode == 1) k<1><<<1, nthreads>>>(d, iters);
      if (mode == 2) k<2><<<1, nthreads>>>(d, iters);
      cudaError_t e = cudaDeviceSynchronize();
      cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
      const char* names[3] = {"ld 128 cols (4 x LDTM.x32, 16 KB/warp)", "st 64 cols (2 x STTM.x32, 8 KB/warp)", "128 dependent ex2"};
      printf("threads=%d %-42s: %.1f clk per iteration per warp (%s)\n", nthreads, names[mode], (double)h[0] / iters,
             cudaGetErrorString(e));
    }
  }
  return 0;
}
