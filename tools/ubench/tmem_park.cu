// Micro-benchmark: Tensor Memory (TMEM) used as a lane-private parking space for a SIMT kernel.
// Each warp stores NREG 32-bit registers per lane with tcgen05.st, loads them back with tcgen05.ld and
// checks them; then times a store+load round trip against the same through shared memory.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tmem_park tmem_park.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>

__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
               :: "r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
                  "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                 "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]) : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

constexpr int WARPS = 8;
constexpr int COLS_PER_WARP = 128;   // 8 warps: 2 per lane quadrant -> 256 of the 512 columns

__global__ void __launch_bounds__(WARPS * 32, 1) k(int iters, unsigned* err, long long* cyc_tmem, long long* cyc_smem) {
  __shared__ uint32_t tbase_s;
  extern __shared__ uint32_t sbuf_raw[];
  uint32_t (*sbuf)[COLS_PER_WARP][33] = reinterpret_cast<uint32_t (*)[COLS_PER_WARP][33]>(sbuf_raw);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" :: "r"((uint32_t)__cvta_generic_to_shared(&tbase_s)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tbase = tbase_s;
  // lane quadrant of this warp (warp % 4) in bits 31:16, column in bits 15:0
  const uint32_t my = tbase + ((uint32_t)(32 * (warp & 3)) << 16) + (uint32_t)((warp >> 2) * COLS_PER_WARP);
  uint32_t r[16], q[16];
  unsigned bad = 0;
  for (int c0 = 0; c0 < COLS_PER_WARP; c0 += 16) {
    for (int i = 0; i < 16; ++i) r[i] = (blockIdx.x << 24) ^ (warp << 16) ^ (lane << 8) ^ (c0 + i);
    tmem_st16(my + c0, r);
  }
  tmem_wait_st();
  for (int c0 = 0; c0 < COLS_PER_WARP; c0 += 16) {
    tmem_ld16(my + c0, q);
    tmem_wait_ld();
    for (int i = 0; i < 16; ++i) bad += (q[i] != ((blockIdx.x << 24) ^ (warp << 16) ^ (lane << 8) ^ (c0 + i)));
  }
  atomicAdd(err, bad);
  // timing: park 128 registers and fetch them back, TMEM vs shared memory
  for (int i = 0; i < 16; ++i) r[i] = lane + i;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    for (int c0 = 0; c0 < COLS_PER_WARP; c0 += 16) tmem_st16(my + c0, r);
    tmem_wait_st();
    for (int c0 = 0; c0 < COLS_PER_WARP; c0 += 16) {
      tmem_ld16(my + c0, q);
      tmem_wait_ld();
      for (int i = 0; i < 16; ++i) r[i] += q[i];
    }
  }
  long long t1 = clock64();
  for (int it = 0; it < iters; ++it) {
    for (int c0 = 0; c0 < COLS_PER_WARP; c0 += 16)
      for (int i = 0; i < 16; ++i) sbuf[warp][c0 + i][lane] = r[i];
    __syncwarp();
    for (int c0 = 0; c0 < COLS_PER_WARP; c0 += 16)
      for (int i = 0; i < 16; ++i) r[i] += sbuf[warp][c0 + i][lane];
    __syncwarp();
  }
  long long t2 = clock64();
  if (threadIdx.x == 0 && blockIdx.x == 0) { *cyc_tmem = t1 - t0; *cyc_smem = t2 - t1; }
  unsigned s = 0; for (int i = 0; i < 16; ++i) s += r[i];
  if (s == 0xdeadbeef) atomicAdd(err, 1);
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" :: "r"(tbase));
}
int main() {
  unsigned* err; long long *ct, *cs;
  cudaMalloc(&err, 4); cudaMalloc(&ct, 8); cudaMalloc(&cs, 8); cudaMemset(err, 0, 4);
  const int iters = 200;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, WARPS * COLS_PER_WARP * 33 * 4);
  k<<<148, WARPS * 32, WARPS * COLS_PER_WARP * 33 * 4>>>(iters, err, ct, cs);
  cudaError_t e = cudaDeviceSynchronize();
  unsigned h; long long ht, hs;
  cudaMemcpy(&h, err, 4, cudaMemcpyDeviceToHost); cudaMemcpy(&ht, ct, 8, cudaMemcpyDeviceToHost); cudaMemcpy(&hs, cs, 8, cudaMemcpyDeviceToHost);
  printf("status %s, mismatches %u\n", cudaGetErrorString(e), h);
  printf("park+fetch of 128 regs/lane, 8 warps/SM: TMEM %.0f cycles/round trip, shared memory %.0f cycles/round trip\n",
         double(ht) / iters, double(hs) / iters);
  return (e != cudaSuccess) || h;
}
