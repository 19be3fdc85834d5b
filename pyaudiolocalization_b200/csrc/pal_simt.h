// pal_simt.h -- the thin SIMT layer every kernel body in this directory is written against.
//
// Compiled by nvcc (sm_100a) it maps 1:1 onto CUDA intrinsics (threadIdx, __syncthreads,
// warp shuffles, mbarrier + cp.async.bulk "TMA" bulk copies).  Compiled by a host C++20
// compiler with -DPAL_EMU it runs the SAME kernel bodies on OS threads (one std::thread per
// CUDA thread, std::barrier for block/warp barriers).  The emulation exists only for the
// CPU test-suite (tests/emu/): there is no GPU in the build container, so kernel logic is
// checked against the oracle on the host before a GPU minute is spent.  The product library
// (libpal_b200.so) never contains the emulation.
#pragma once

#include <cstddef>
#include <cstdint>
#include <cmath>

#if defined(__CUDACC__) && !defined(PAL_EMU)
#define PAL_GPU 1
#define PAL_HD __host__ __device__ __forceinline__
#define PAL_DEV __device__ __forceinline__
#else
#define PAL_GPU 0
#define PAL_HD inline
#define PAL_DEV inline
#include <atomic>
#include <barrier>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <thread>
#include <vector>
#endif

namespace pal {

template <typename T> struct alignas(2 * sizeof(T)) cpx { T x, y; };
using cpxf = cpx<float>;
using cpxd = cpx<double>;

struct alignas(8) mbar_t { unsigned long long v; };

#if PAL_GPU
// ------------------------------------------------------------------ device
namespace simt {
PAL_DEV int tid() { return threadIdx.x; }
PAL_DEV int nthreads() { return blockDim.x; }
PAL_DEV int bid() { return blockIdx.x; }
PAL_DEV int nblocks() { return gridDim.x; }
PAL_DEV int lane() { return threadIdx.x & 31; }
PAL_DEV int warp() { return threadIdx.x >> 5; }
PAL_DEV void sync_block() { __syncthreads(); }
PAL_DEV void sync_warp() { __syncwarp(); }
template <class T> PAL_DEV T shfl_xor(T v, int m) { return __shfl_xor_sync(0xffffffffu, v, m); }
template <class T> PAL_DEV T shfl(T v, int src) { return __shfl_sync(0xffffffffu, v, src); }
PAL_DEV unsigned ballot(bool p) { return __ballot_sync(0xffffffffu, p); }

// mbarrier + 1-D TMA bulk copy (cp.async.bulk, SASS: UBLKCP)
PAL_DEV unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
PAL_DEV void mbar_init(mbar_t* b, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(count));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
PAL_DEV void mbar_expect_tx(mbar_t* b, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes)
               : "memory");
}
PAL_DEV void bulk_g2s(void* dst_smem, const void* src_gmem, unsigned bytes, mbar_t* b) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
          smem_u32(dst_smem)),
      "l"(src_gmem), "r"(bytes), "r"(smem_u32(b))
      : "memory");
}
// order earlier generic-proxy accesses of shared memory before a following bulk (async-proxy) copy
PAL_DEV void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
PAL_DEV void mbar_wait(mbar_t* b, unsigned parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t}" ::"r"(smem_u32(b)),
      "r"(parity)
      : "memory");
}
}  // namespace simt

// ---- packed pair of fp32 values (Blackwell FFMA2 / FADD2 / FMUL2: fma.rn.f32x2 & co, sm_100+).
// One issue slot does two fp32 operations; ptxas folds half swaps / per-half negations of an
// operand (f2_make(-hi, lo) ...) and scalar broadcasts into operand modifiers, so a complex
// number carried as (re, im) in one aligned register pair costs nothing extra for *i, conj.
struct f2 { unsigned long long v; };
PAL_DEV f2 f2_make(float lo, float hi) { f2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r.v) : "f"(lo), "f"(hi)); return r; }
PAL_DEV float f2_lo(f2 a) { float lo, hi; asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(a.v)); return lo; }
PAL_DEV float f2_hi(f2 a) { float lo, hi; asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(a.v)); return hi; }
PAL_DEV f2 f2_add(f2 a, f2 b) { f2 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v)); return r; }
PAL_DEV f2 f2_sub(f2 a, f2 b) { f2 r; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v)); return r; }
PAL_DEV f2 f2_mul(f2 a, f2 b) { f2 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v)); return r; }
PAL_DEV f2 f2_fma(f2 a, f2 b, f2 c) { f2 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r.v) : "l"(a.v), "l"(b.v), "l"(c.v)); return r; }

// ---- Tensor Memory (TMEM, 512 columns x 128 lanes x 32 bit per SM) as LANE-PRIVATE scratch of a SIMT
// kernel: a warp reads / writes the 32 TMEM lanes of its quadrant (warp % 4), lane i <-> thread i,
// N consecutive 32-bit columns per instruction (tcgen05.ld / tcgen05.st .32x32b, SASS LDTM / STTM).
// Used to park register tiles so that more warps fit on an SM (see pair4095_tmem_body).
namespace simt {
PAL_DEV void tmem_alloc512(unsigned* smem_slot) {   // one full warp calls this
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(smem_slot)) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
PAL_DEV void tmem_dealloc512(unsigned base) {       // one full warp calls this
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(base) : "memory");
}
PAL_DEV void tmem_fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
PAL_DEV void tmem_fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
PAL_DEV void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
PAL_DEV void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// address of column `col` in the quadrant of warp `warp`
PAL_DEV unsigned tmem_addr(unsigned base, int warp, int col) { return base + ((unsigned)(32 * (warp & 3)) << 16) + (unsigned)col; }
}  // namespace simt
PAL_DEV unsigned f2_bits_lo(f2 a) { return (unsigned)(a.v & 0xffffffffull); }
PAL_DEV unsigned f2_bits_hi(f2 a) { return (unsigned)(a.v >> 32); }
PAL_DEV f2 f2_from_bits(unsigned lo, unsigned hi) { f2 r; r.v = ((unsigned long long)hi << 32) | lo; return r; }
// park / fetch N complex values (2N columns)
PAL_DEV void tmem_st(unsigned ta, const f2 (&v)[1]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x2.b32 [%0], {%1,%2};" ::"r"(ta), "r"(f2_bits_lo(v[0])), "r"(f2_bits_hi(v[0])) : "memory");
}
PAL_DEV void tmem_ld(unsigned ta, f2 (&v)[1]) {
  unsigned r0, r1;
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0,%1}, [%2];" : "=r"(r0), "=r"(r1) : "r"(ta) : "memory");
  v[0] = f2_from_bits(r0, r1);
}
PAL_DEV void tmem_ld(unsigned ta, f2 (&v)[4]) {
  unsigned r[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(ta) : "memory");
#pragma unroll
  for (int i = 0; i < 4; ++i) v[i] = f2_from_bits(r[2 * i], r[2 * i + 1]);
}
PAL_DEV void tmem_ld(unsigned ta, f2 (&v)[8]) {
  unsigned r[16];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                 "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]) : "r"(ta) : "memory");
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = f2_from_bits(r[2 * i], r[2 * i + 1]);
}

PAL_DEV float fma_(float a, float b, float c) { return fmaf(a, b, c); }
PAL_DEV double fma_(double a, double b, double c) { return fma(a, b, c); }
PAL_DEV float sqrt_(float a) { return sqrtf(a); }
PAL_DEV double sqrt_(double a) { return sqrt(a); }
PAL_DEV float abs_(float a) { return fabsf(a); }
PAL_DEV double abs_(double a) { return fabs(a); }
PAL_DEV float max_(float a, float b) { return fmaxf(a, b); }
PAL_DEV double max_(double a, double b) { return fmax(a, b); }
PAL_DEV float min_(float a, float b) { return fminf(a, b); }
PAL_DEV double min_(double a, double b) { return fmin(a, b); }

#else
// ------------------------------------------------------------------ host emulation
}  // namespace pal
struct alignas(16) float4 { float x, y, z, w; };
namespace pal {
namespace simt {
struct WarpCtx {
  std::unique_ptr<std::barrier<>> bar;
  unsigned long long slot[32];
  int width;
};
struct BlockCtx {
  std::unique_ptr<std::barrier<>> bar;
  std::vector<WarpCtx> warps;
  int nthreads, bid, nblocks;
};
struct ThreadCtx {
  BlockCtx* blk = nullptr;
  int tid = 0;
};
inline thread_local ThreadCtx tctx;

inline int tid() { return tctx.tid; }
inline int nthreads() { return tctx.blk->nthreads; }
inline int bid() { return tctx.blk->bid; }
inline int nblocks() { return tctx.blk->nblocks; }
inline int lane() { return tctx.tid & 31; }
inline int warp() { return tctx.tid >> 5; }
inline void sync_block() { tctx.blk->bar->arrive_and_wait(); }
inline void sync_warp() { tctx.blk->warps[warp()].bar->arrive_and_wait(); }
template <class T> inline T shfl(T v, int src) {
  static_assert(sizeof(T) <= 8, "shfl payload");
  WarpCtx& w = tctx.blk->warps[warp()];
  unsigned long long raw = 0;
  std::memcpy(&raw, &v, sizeof(T));
  w.slot[lane()] = raw;
  w.bar->arrive_and_wait();
  unsigned long long got = w.slot[(src & 31) < w.width ? (src & 31) : lane()];
  w.bar->arrive_and_wait();
  T out;
  std::memcpy(&out, &got, sizeof(T));
  return out;
}
template <class T> inline T shfl_xor(T v, int m) { return shfl(v, lane() ^ m); }
inline unsigned ballot(bool p) {
  unsigned bit = p ? (1u << lane()) : 0u;
  unsigned acc = 0;
  WarpCtx& w = tctx.blk->warps[warp()];
  w.slot[lane()] = bit;
  w.bar->arrive_and_wait();
  for (int i = 0; i < w.width; ++i) acc |= (unsigned)w.slot[i];
  w.bar->arrive_and_wait();
  return acc;
}

// mbarrier emulation: v = (completed_phases << 32) | pending_bytes ; single producer thread.
inline std::atomic<unsigned long long>* mb(mbar_t* b) {
  return reinterpret_cast<std::atomic<unsigned long long>*>(&b->v);
}
inline void mbar_init(mbar_t* b, int) { mb(b)->store(0); }
inline void mbar_expect_tx(mbar_t* b, unsigned bytes) { mb(b)->fetch_add(bytes); }
inline void bulk_g2s(void* dst, const void* src, unsigned bytes, mbar_t* b) {
  if ((reinterpret_cast<uintptr_t>(dst) | reinterpret_cast<uintptr_t>(src) | bytes) & 15u) std::abort();
  std::memcpy(dst, src, bytes);
  unsigned long long before = mb(b)->fetch_sub(bytes);
  if ((before & 0xffffffffull) == bytes) mb(b)->fetch_add(1ull << 32);  // phase complete
}
inline void fence_async_smem() {}
inline void mbar_wait(mbar_t* b, unsigned parity) {
  while (((mb(b)->load(std::memory_order_acquire) >> 32) & 1ull) == parity) std::this_thread::yield();
}

// Run `body(smem)` as a grid of `grid` blocks of `block` threads (blocks sequentially).
template <class F> inline void launch(int grid, int block, size_t smem_bytes, F body) {
  for (int b = 0; b < grid; ++b) {
    BlockCtx blk;
    blk.nthreads = block;
    blk.bid = b;
    blk.nblocks = grid;
    blk.bar = std::make_unique<std::barrier<>>(block);
    int nw = (block + 31) / 32;
    blk.warps.resize(nw);
    for (int w = 0; w < nw; ++w) {
      int width = (w == nw - 1) ? block - 32 * w : 32;
      blk.warps[w].width = width;
      blk.warps[w].bar = std::make_unique<std::barrier<>>(width);
    }
    void* smem = nullptr;
    if (posix_memalign(&smem, 1024, smem_bytes ? smem_bytes : 16)) std::abort();
    std::memset(smem, 0xA5, smem_bytes);  // poison: catches reads of unwritten shared memory
    std::vector<std::thread> th;
    th.reserve(block);
    for (int t = 0; t < block; ++t)
      th.emplace_back([&, t] {
        tctx.blk = &blk;
        tctx.tid = t;
        body(static_cast<char*>(smem));
      });
    for (auto& x : th) x.join();
    std::free(smem);
  }
}
}  // namespace simt

struct alignas(8) f2 { float lo, hi; };
inline f2 f2_make(float lo, float hi) { return f2{lo, hi}; }
inline float f2_lo(f2 a) { return a.lo; }
inline float f2_hi(f2 a) { return a.hi; }
inline f2 f2_add(f2 a, f2 b) { return f2{a.lo + b.lo, a.hi + b.hi}; }
inline f2 f2_sub(f2 a, f2 b) { return f2{a.lo - b.lo, a.hi - b.hi}; }
inline f2 f2_mul(f2 a, f2 b) { return f2{a.lo * b.lo, a.hi * b.hi}; }
inline f2 f2_fma(f2 a, f2 b, f2 c) { return f2{std::fmaf(a.lo, b.lo, c.lo), std::fmaf(a.hi, b.hi, c.hi)}; }

// TMEM emulation: 512 lane-private columns per thread (thread_local), same call surface
namespace simt {
inline thread_local unsigned tmem_cols[512];
inline void tmem_alloc512(unsigned* smem_slot) { if (lane() == 0) *smem_slot = 0u; }
inline void tmem_dealloc512(unsigned) {}
inline void tmem_fence_before_sync() {}
inline void tmem_fence_after_sync() {}
inline void tmem_wait_ld() {}
inline void tmem_wait_st() {}
inline unsigned tmem_addr(unsigned base, int, int col) { return base + (unsigned)col; }
}  // namespace simt
template <int N> inline void tmem_st(unsigned ta, const f2 (&v)[N]) {
  for (int i = 0; i < N; ++i) { std::memcpy(&simt::tmem_cols[ta + 2 * i], &v[i].lo, 4); std::memcpy(&simt::tmem_cols[ta + 2 * i + 1], &v[i].hi, 4); }
}
template <int N> inline void tmem_ld(unsigned ta, f2 (&v)[N]) {
  for (int i = 0; i < N; ++i) { std::memcpy(&v[i].lo, &simt::tmem_cols[ta + 2 * i], 4); std::memcpy(&v[i].hi, &simt::tmem_cols[ta + 2 * i + 1], 4); }
}

inline float fma_(float a, float b, float c) { return std::fmaf(a, b, c); }
inline double fma_(double a, double b, double c) { return std::fma(a, b, c); }
inline float sqrt_(float a) { return std::sqrt(a); }
inline double sqrt_(double a) { return std::sqrt(a); }
inline float abs_(float a) { return std::fabs(a); }
inline double abs_(double a) { return std::fabs(a); }
inline float max_(float a, float b) { return std::fmax(a, b); }
inline double max_(double a, double b) { return std::fmax(a, b); }
inline float min_(float a, float b) { return std::fmin(a, b); }
inline double min_(double a, double b) { return std::fmin(a, b); }
#endif

// helpers common to both builds: complex numbers as (re, im) pairs
PAL_DEV f2 f2_bcast(float c) { return f2_make(c, c); }
PAL_DEV f2 f2_muli(f2 a) { return f2_make(-f2_hi(a), f2_lo(a)); }     // i * a
PAL_DEV f2 f2_conj(f2 a) { return f2_make(f2_lo(a), -f2_hi(a)); }
PAL_DEV f2 f2_swap(f2 a) { return f2_make(f2_hi(a), f2_lo(a)); }

}  // namespace pal
