// Micro-benchmark: is ex2.approx.ftz.bf16x2 (two exponentials per lane per MUFU instruction) twice the
// element rate of ex2.approx.ftz.f32 on B200?  Also times the candidate packed softmax inner loop.
#include <cstdio>
#include <cstdint>
#include <cuda_bf16.h>
__device__ __forceinline__ float ex2f(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ uint32_t ex2h2(uint32_t x) { uint32_t y; asm("ex2.approx.ftz.bf16x2 %0, %1;" : "=r"(y) : "r"(x)); return y; }
__device__ __forceinline__ uint32_t pack(float lo, float hi) { uint32_t r; asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo)); return r; }
__device__ __forceinline__ uint32_t fma2(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; asm("fma.rn.bf16x2 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }

template <int MODE>
__global__ void __launch_bounds__(1024, 1) k(long long* out, float* sink, int iters, float sc, float nm) {
  float s[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) s[i] = threadIdx.x * 0.001f - i * 0.1f;
  float acc = 0.f; uint32_t pacc = 0;
  const uint32_t sc2 = pack(sc, sc), nm2 = pack(nm, nm);
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    float r0 = 0.f, r1 = 0.f;
#pragma unroll
    for (int i = 0; i < 32; i += 2) {
      if (MODE == 0) {            // fp32 path: 2x (ffma, ex2, fadd) + pack
        const float e0 = ex2f(fmaf(s[i], sc, nm)), e1 = ex2f(fmaf(s[i + 1], sc, nm));
        r0 += e0; r1 += e1; pacc ^= pack(e0, e1);
      } else if (MODE == 1) {     // packed: pack, fma.bf16x2, ex2.bf16x2, unpack-sum in fp32
        const uint32_t e = ex2h2(fma2(pack(s[i], s[i + 1]), sc2, nm2));
        r0 += __uint_as_float(e << 16); r1 += __uint_as_float(e & 0xffff0000u); pacc ^= e;
      } else if (MODE == 2) {     // packed, x computed in fp32 then packed
        const uint32_t e = ex2h2(pack(fmaf(s[i], sc, nm), fmaf(s[i + 1], sc, nm)));
        r0 += __uint_as_float(e << 16); r1 += __uint_as_float(e & 0xffff0000u); pacc ^= e;
      } else {                    // raw MUFU rate: ex2.bf16x2 only
        pacc ^= ex2h2(__float_as_uint(s[i]));
        pacc ^= ex2h2(__float_as_uint(s[i + 1]));
      }
    }
    acc += r0 + r1;
#pragma unroll
    for (int i = 0; i < 32; ++i) s[i] += acc * 1e-9f;
  }
  long long t1 = clock64();
  if ((threadIdx.x & 31) == 0) out[threadIdx.x >> 5] = t1 - t0;
  if (acc == 1234.5f) sink[0] = acc + pacc;
}
int main() {
  long long* d; float* sink; cudaMalloc(&d, 64 * 8); cudaMalloc(&sink, 4);
  long long h[64]; const int iters = 4000;
  const char* names[4] = {"fp32: ffma+ex2.f32+fadd, pack", "bf16x2: pack,fma2,ex2.bf16x2,unpack-sum", "bf16x2: ffma fp32, pack, ex2.bf16x2, unpack-sum", "raw ex2.bf16x2 x32 (64 results)"};
  for (int nthreads : {128, 512}) for (int mode = 0; mode < 4; ++mode) {
    if (mode == 0) k<0><<<1, nthreads>>>(d, sink, iters, 0.127f, -3.f);
    if (mode == 1) k<1><<<1, nthreads>>>(d, sink, iters, 0.127f, -3.f);
    if (mode == 2) k<2><<<1, nthreads>>>(d, sink, iters, 0.127f, -3.f);
    if (mode == 3) k<3><<<1, nthreads>>>(d, sink, iters, 0.127f, -3.f);
    cudaError_t e = cudaDeviceSynchronize();
    cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    printf("warps/SMSP=%d %-48s: %7.1f clk per 32 elements per warp -> %5.2f clk/elem/SMSP (%s)\n", nthreads / 128, names[mode],
           (double)h[0] / iters, (double)h[0] / iters / 32.0 / (nthreads / 128), cudaGetErrorString(e));
  }
  return 0;
}
