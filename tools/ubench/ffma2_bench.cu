// Micro-benchmark: issue rate of scalar FFMA vs packed FFMA2 / FADD2 (fma.rn.f32x2, sm_100a)
// and how they mix with shared-memory and integer instructions.  Build:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma2_bench ffma2_bench.cu
#include <cuda_runtime.h>
#include <cstdio>
typedef unsigned long long u64;
__device__ __forceinline__ u64 pk(float a, float b) { u64 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ u64 ffma2(u64 a, u64 b, u64 c) { u64 d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ u64 fadd2(u64 a, u64 b) { u64 d; asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ float ffma1(float a, float b, float c) { float d; asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c)); return d; }

constexpr int ITERS = 4096;
// mode 0: 8 independent scalar FFMA chains   mode 1: 8 independent FFMA2 chains
// mode 2: FFMA2 + FADD2 alternating          mode 3: FFMA2 : LDS 2:1     mode 4: FFMA : LDS 2:1
// mode 5: FFMA2 : IADD 1:1                   mode 6: scalar FFMA : IADD 1:1   mode 7: FFMA2 : scalar FFMA 1:1
template <int MODE> __global__ void __launch_bounds__(1024) kern(float* out, float seed) {
  __shared__ float sm[1024];
  sm[threadIdx.x] = seed;
  __syncthreads();
  float a[8]; u64 p[8]; int ia[8];
  for (int i = 0; i < 8; ++i) { a[i] = seed + i; p[i] = pk(seed + i, seed - i); ia[i] = threadIdx.x + i; }
  const float m = 0.999f; const u64 mp = pk(m, m); const u64 cp = pk(seed, seed);
  float ld = 0.f;
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MODE == 0) a[i] = ffma1(a[i], m, seed);
      if (MODE == 1) p[i] = ffma2(p[i], mp, cp);
      if (MODE == 2) { if (i & 1) p[i] = ffma2(p[i], mp, cp); else p[i] = fadd2(p[i], cp); }
      if (MODE == 3) { p[i] = ffma2(p[i], mp, cp); if (i & 1) ld += sm[(threadIdx.x + i * 32 + it) & 1023]; }
      if (MODE == 4) { a[i] = ffma1(a[i], m, seed); if (i & 1) ld += sm[(threadIdx.x + i * 32 + it) & 1023]; }
      if (MODE == 5) { p[i] = ffma2(p[i], mp, cp); asm volatile("add.s32 %0, %0, %1;" : "+r"(ia[i]) : "r"(it)); }
      if (MODE == 6) { a[i] = ffma1(a[i], m, seed); asm volatile("add.s32 %0, %0, %1;" : "+r"(ia[i]) : "r"(it)); }
      if (MODE == 7) { p[i] = ffma2(p[i], mp, cp); a[i] = ffma1(a[i], m, seed); }
    }
  }
  float s = ld;
  for (int i = 0; i < 8; ++i) { s += a[i] + float(p[i] & 0xffff) + ia[i]; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int MODE> void run(const char* name, int threads, double fp_per_iter, double inst_per_iter) {
  int dev, sms; cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, dev);
  float* out; cudaMalloc(&out, sizeof(float) * sms * 1024);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  kern<MODE><<<sms, threads>>>(out, 1.0f);
  cudaDeviceSynchronize();
  float best = 1e9;
  for (int r = 0; r < 5; ++r) {
    cudaEventRecord(e0); kern<MODE><<<sms, threads>>>(out, 1.0f); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
  }
  const double cycles = best * 1e-3 * clk * 1e3;    // at max clock (kHz)
  const double warps = threads / 32.0;
  const double winst = warps * ITERS * inst_per_iter;      // warp instructions per SM
  const double flane = warps * 32 * ITERS * fp_per_iter;   // fp32 FMA lanes-ops per SM
  printf("%-28s threads/SM=%4d  %.3f ms  warp-inst/clk/SM=%.2f  fp32-ops/clk/SM=%.1f\n", name, threads, best,
         winst / cycles, flane / cycles);
  cudaFree(out);
}
int main() {
  for (int th : {128, 256, 512, 1024}) {
    run<0>("FFMA", th, 8, 8);
    run<1>("FFMA2", th, 16, 8);
    run<2>("FFMA2+FADD2", th, 16, 8);
    run<3>("FFMA2:LDS 2:1", th, 16, 12);
    run<4>("FFMA:LDS 2:1", th, 8, 12);
    run<5>("FFMA2:IADD 1:1", th, 16, 16);
    run<6>("FFMA:IADD 1:1", th, 8, 16);
    run<7>("FFMA2:FFMA 1:1", th, 24, 16);
  }
  return 0;
}
