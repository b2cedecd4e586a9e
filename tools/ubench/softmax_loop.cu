// Micro-benchmark: candidate softmax inner loops for the forward kernel, per warp-element per SM sub-partition.
//   mode 0  scalar      : FFMA + MUFU.EX2 + FADD + 1/2 F2FP            (what the kernel did in round 1)
//   mode 1  packed      : 1/2 FFMA2 + MUFU.EX2 + 1/2 FADD2 + 1/2 F2FP
//   mode 2-5 packed+emu : as 1, with 2 / 3 / 4 / 8 of every 8 exponentials evaluated by a packed polynomial
//   mode 6  max pass    : FMNMX (1 per element)      mode 7: FMNMX3 (1/2 per element)
//   mode 8  raw FFMA2   : 32 independent FFMA2 (64 fp32 FMAs)      mode 9: 64 independent FFMA
#include <cstdio>
#include <cstdint>
__device__ __forceinline__ float ex2f(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ uint32_t pack(float lo, float hi) { uint32_t r; asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo)); return r; }
__device__ __forceinline__ float max3(float a, float b, float c) { float d; asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c)); return d; }
__device__ __forceinline__ uint64_t pk2(float lo, float hi) { uint64_t r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void up2(uint64_t v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) { uint64_t r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ uint64_t add2(uint64_t a, uint64_t b) { uint64_t r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ uint64_t sub2(uint64_t a, uint64_t b) { uint64_t r; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
// 2^x for a pair, FMA/ALU pipes only (x clamped at -126)
__device__ __forceinline__ void ex2_poly2(uint64_t x2, float& e0, float& e1) {
  float x0, x1; up2(x2, x0, x1);
  x0 = fmaxf(x0, -126.f); x1 = fmaxf(x1, -126.f);
  x2 = pk2(x0, x1);
  const uint64_t magic = pk2(12582912.f, 12582912.f);
  const uint64_t t2 = add2(x2, magic);
  const uint64_t f2 = sub2(x2, sub2(t2, magic));
  uint64_t p2 = fma2(pk2(0.0551716685f, 0.0551716685f), f2, pk2(0.2426111251f, 0.2426111251f));
  p2 = fma2(p2, f2, pk2(0.6932609677f, 0.6932609677f));
  p2 = fma2(p2, f2, pk2(0.9999280572f, 0.9999280572f));
  float p0, p1, t0, t1; up2(p2, p0, p1); up2(t2, t0, t1);
  e0 = __int_as_float(__float_as_int(p0) + (__float_as_int(t0) << 23));
  e1 = __int_as_float(__float_as_int(p1) + (__float_as_int(t1) << 23));
}

template <int MODE>
__global__ void __launch_bounds__(1024, 1) k(long long* out, float* sink, int iters, float sc, float nm) {
  float s[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) s[i] = threadIdx.x * 0.001f - i * 0.05f;
  float acc = 0.f; uint32_t pacc = 0;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    const float nmi = nm + it * 1e-7f;   // loop-variant so nothing is hoisted
    if (MODE == 0) {
      float r0 = 0.f, r1 = 0.f, r2 = 0.f, r3 = 0.f;
#pragma unroll
      for (int i = 0; i < 32; i += 4) {
        const float e0 = ex2f(fmaf(s[i], sc, nmi)), e1 = ex2f(fmaf(s[i + 1], sc, nmi));
        const float e2 = ex2f(fmaf(s[i + 2], sc, nmi)), e3 = ex2f(fmaf(s[i + 3], sc, nmi));
        r0 += e0; r1 += e1; r2 += e2; r3 += e3; pacc ^= pack(e0, e1) ^ pack(e2, e3);
      }
      acc += (r0 + r1) + (r2 + r3);
    } else if (MODE >= 1 && MODE <= 5) {
      constexpr int EMU = (MODE == 1) ? 0 : (MODE == 2) ? 2 : (MODE == 3) ? 3 : (MODE == 4) ? 4 : 8;
      const uint64_t sc2 = pk2(sc, sc), nm2 = pk2(nmi, nmi);
      uint64_t ra = pk2(0.f, 0.f), rb = pk2(0.f, 0.f);
#pragma unroll
      for (int i = 0; i < 32; i += 8) {
        float e[8];
#pragma unroll
        for (int t = 0; t < 8; t += 2) {
          const uint64_t x2 = fma2(pk2(s[i + t], s[i + t + 1]), sc2, nm2);
          // pairs emulated: EMU=2 -> pair 3; 4 -> pairs 1,3; 3 -> pair 3 + half of pair 1 (approximated as pair 3 and every other block pair 1)
          const bool emu = (EMU == 8) || (EMU >= 2 && t == 6) || (EMU >= 4 && t == 2) || (EMU == 3 && t == 2 && (i & 8));
          if (emu) ex2_poly2(x2, e[t], e[t + 1]);
          else { float x0, x1; up2(x2, x0, x1); e[t] = ex2f(x0); e[t + 1] = ex2f(x1); }
        }
        ra = add2(ra, add2(pk2(e[0], e[1]), pk2(e[4], e[5])));
        rb = add2(rb, add2(pk2(e[2], e[3]), pk2(e[6], e[7])));
        pacc ^= pack(e[0], e[1]) ^ pack(e[2], e[3]) ^ pack(e[4], e[5]) ^ pack(e[6], e[7]);
      }
      float a0, a1; up2(add2(ra, rb), a0, a1);
      acc += a0 + a1;
    } else if (MODE == 6) {
      float a0 = nmi, a1 = nmi, a2 = nmi, a3 = nmi;
#pragma unroll
      for (int i = 0; i < 32; i += 4) { a0 = fmaxf(a0, s[i] * sc); a1 = fmaxf(a1, s[i + 1]); a2 = fmaxf(a2, s[i + 2]); a3 = fmaxf(a3, s[i + 3]); }
      acc += fmaxf(fmaxf(a0, a1), fmaxf(a2, a3));
    } else if (MODE == 7) {
      float a0 = nmi, a1 = nmi;
#pragma unroll
      for (int i = 0; i < 32; i += 4) { a0 = max3(a0, s[i] * sc, s[i + 1]); a1 = max3(a1, s[i + 2], s[i + 3]); }
      acc += fmaxf(a0, a1);
    } else if (MODE == 8) {
      const uint64_t sc2 = pk2(sc, sc), nm2 = pk2(nmi, nmi);
      uint64_t r[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) r[i] = pk2(s[2 * i], s[2 * i + 1]);
#pragma unroll
      for (int rep = 0; rep < 2; ++rep)
#pragma unroll
        for (int i = 0; i < 16; ++i) r[i] = fma2(r[i], sc2, nm2);
#pragma unroll
      for (int i = 0; i < 16; ++i) { float a, b; up2(r[i], a, b); pacc ^= __float_as_uint(a) ^ __float_as_uint(b); }
    } else {
      float r[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) r[i] = s[i];
#pragma unroll
      for (int rep = 0; rep < 2; ++rep)
#pragma unroll
        for (int i = 0; i < 32; ++i) r[i] = fmaf(r[i], sc, nmi);
#pragma unroll
      for (int i = 0; i < 32; ++i) pacc ^= __float_as_uint(r[i]);
    }
  }
  long long t1 = clock64();
  if ((threadIdx.x & 31) == 0) out[threadIdx.x >> 5] = t1 - t0;
  if (acc == 1234.5f || pacc == 0x12345u) sink[0] = acc + pacc;
}

template <int MODE>
void run(const char* name, long long* d, float* sink) {
  long long h[64];
  const int iters = 4000;
  for (int nthreads : {128, 256, 512}) {
    k<MODE><<<1, nthreads>>>(d, sink, iters, 0.127f, -3.f);
    cudaError_t e = cudaDeviceSynchronize();
    cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    printf("warps/SMSP=%d %-34s: %7.1f clk per 32-element block per warp -> %5.2f clk/elem/SMSP (%s)\n", nthreads / 128, name,
           (double)h[0] / iters, (double)h[0] / iters / 32.0 / (nthreads / 128), cudaGetErrorString(e));
  }
}
int main() {
  long long* d; float* sink; cudaMalloc(&d, 64 * 8); cudaMalloc(&sink, 4);
  run<0>("scalar ffma+ex2+fadd+pack", d, sink);
  run<1>("packed ffma2+ex2+fadd2+pack", d, sink);
  run<2>("packed, 2/8 poly", d, sink);
  run<3>("packed, 3/8 poly", d, sink);
  run<4>("packed, 4/8 poly", d, sink);
  run<5>("packed, 8/8 poly", d, sink);
  run<6>("max pass FMNMX", d, sink);
  run<7>("max pass FMNMX3", d, sink);
  run<8>("raw FFMA2 x32 (=64 fma, 2 elem-units)", d, sink);
  run<9>("raw FFMA x64 (2 elem-units)", d, sink);
  return 0;
}
