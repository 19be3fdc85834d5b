// pal_generic_host.cuh -- host-side orchestration of the arbitrary-length GCC-PHAT path
// (Bluestein, pal_bluestein.cuh).  Included by pal_capi.cu only.
#pragma once
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <cstdlib>
#include <map>
#include <mutex>
#include <type_traits>

#include "pal_bluestein.cuh"
#include "pal_fft2.cuh"
#include "pal_winpick.cuh"
#include "pal_sync.cuh"

namespace palhost {
using namespace pal;

extern std::atomic<unsigned long long>* g_launch_counter;

constexpr int kGT = 256;  // threads per block of every generic kernel
template <typename T> struct ColTile { static constexpr int TC = 16; static constexpr int TR = 16; };
template <> struct ColTile<double> { static constexpr int TC = 8; static constexpr int TR = 8; };

#ifndef PAL_GEN_MINBLOCKS
#define PAL_GEN_MINBLOCKS 4
#endif
// resident blocks per SM the float32 FFT passes are compiled for (register cap); float64 is left to the compiler
template <typename T> struct MinBlocks { static constexpr int V = PAL_GEN_MINBLOCKS; };
template <> struct MinBlocks<double> { static constexpr int V = 1; };

template <typename T> __global__ void __launch_bounds__(kGT) k_blue_init(BluePlan p, cpx<T>* chirp, cpx<T>* tw1,
                                                                       cpx<T>* tw2, cpx<T>* twM) {
  blue_init_tables_body<T>(p, chirp, tw1, tw2, twM);
}
template <typename T, class Loader>
__global__ void __launch_bounds__(kGT, MinBlocks<T>::V) k_colpass_fwd(BluePlan p, BlueTables<T> tb, Loader ld, long long n_tr,
                                                     const int* n_tr_dev, cpx<T>* buf) {
  extern __shared__ __align__(128) char smem[];
  colpass_fwd_body<T, kGT, ColTile<T>::TC, Loader>(p, tb, ld, n_tr_dev ? (long long)*n_tr_dev * n_tr : n_tr, buf, smem);
}
template <typename T, bool CONV, bool CONJ, int TRV = ColTile<T>::TR>
__global__ void __launch_bounds__(kGT, MinBlocks<T>::V) k_rowpass(BluePlan p, BlueTables<T> tb, long long n_tr, const int* n_tr_dev,
                                                 cpx<T>* buf) {
  extern __shared__ __align__(128) char smem[];
  rowpass_body<T, kGT, TRV, CONV, CONJ>(p, tb, n_tr_dev ? (long long)*n_tr_dev * n_tr : n_tr, buf, smem);
}
template <typename T, class Storer>
__global__ void __launch_bounds__(kGT, MinBlocks<T>::V) k_colpass_inv(BluePlan p, BlueTables<T> tb, Storer st, long long n_tr,
                                                     const int* n_tr_dev, const cpx<T>* buf) {
  extern __shared__ __align__(128) char smem[];
  colpass_inv_body<T, kGT, ColTile<T>::TC, Storer>(p, tb, st, n_tr_dev ? (long long)*n_tr_dev * n_tr : n_tr, buf, smem);
}
template <typename T>
__global__ void __launch_bounds__(kGT) k_pick_rows(const T* corr, int n, int c0, long long n_rows, const int* n_rows_dev,
                                                   const int* item_list, int win_half, int dist, int method, float mult,
                                                   int num_peaks, float eps, unsigned char* pkmap_ws, int* k_idx,
                                                   int* k_count, float* peak, float* gmax, unsigned* flags,
                                                   unsigned extra_flag, unsigned keep_mask, float* corr_out) {
  extern __shared__ __align__(128) char smem[];
  pick_rows_body<T, kGT>(corr, n, c0, n_rows_dev ? (long long)*n_rows_dev : n_rows, item_list, win_half, dist, method,
                         mult, num_peaks, eps, pkmap_ws, k_idx, k_count, peak, gmax, flags, extra_flag, keep_mask, corr_out, smem);
}
__global__ void __launch_bounds__(kGT) k_win_pick(const float* win, const float* pmax, int tiles, long long n_rows, WinGeom g,
                                                  long long item0, int* k_idx, int* k_count, float* peak, float* gmax,
                                                  unsigned* flags, unsigned extra_flag, WhitenRef wr) {
  win_pick_rows_body<kGT>(win, pmax, tiles, n_rows, g, item0, k_idx, k_count, peak, gmax, flags, extra_flag, wr);
}
__global__ void __launch_bounds__(kGT) k_whiten_unpack(const cpxf* Z, int n, long long n_packed, int Mics, int CP, const float* scales,
                                                       long long row_base, long long local_row_base, cpxf* U, float* hq) {
  __shared__ float sh[4 * kGT / 32];
  whiten_unpack_body<kGT>(Z, n, n_packed, Mics, CP, scales, row_base, local_row_base, U, hq, reinterpret_cast<char*>(sh));
}
template <typename TS>
__global__ void __launch_bounds__(kGT) k_row_scales(const TS* sig, long long n_rows, long long ld, int len_even, int len_odd,
                                                    float* scales) {
  __shared__ float sh[kGT / 32];
  row_scale_body<kGT, TS>(sig, n_rows, ld, len_even, len_odd, scales, reinterpret_cast<char*>(sh));
}
// sweeps that pick from full rows (k_pick_rows: several peaks, correlation rows wanted, short transforms) carry no
// per-row margin: rows whose rounding-noise bound is not small against the tie margin go to the float64 sweep outright
__global__ void k_noise_flags(WhitenRef wr, long long n_items, int n, float eps, unsigned* flags) {
  const long long it = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (it < n_items && !(pair_bound(wr, it, n) <= 0.5f * eps)) flags[it] |= PAL_FLAG_NEAR_TIE;
}
// device-side chunking of a list whose length only the device knows: round k of `chunk` items holds
// items[k] = clamp(count - k * chunk, 0, chunk) items = packed[k] = ceil(items[k] / 2) packed inverse transforms
__global__ void k_chunk_counts(const int* count, int chunk, int rounds, int* items, int* packed) {
  const int k = threadIdx.x;          // launched with >= kMaxDevRounds threads
  if (k < rounds) {
    long long left = (long long)*count - (long long)k * chunk;
    left = left < 0 ? 0 : (left > chunk ? chunk : left);
    items[k] = int(left);
    packed[k] = int((left + 1) / 2);
  }
}
constexpr int kMaxDevRounds = 96;    // rounds of the device-counted float64 sweep (two int arrays next to the counter)

// flagged item -> its two channel rows (for the float64 re-evaluation)
__global__ void k_rows_of_items(const int* item_list, const int* count, const int* pairs, int Mics, int P, int* rows) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < *count) {
    const int item = item_list[i];
    const int f = item / P, p = item % P;
    rows[2 * i] = f * Mics + pairs[2 * p];
    rows[2 * i + 1] = f * Mics + pairs[2 * p + 1];
  }
}

inline size_t al(size_t v) { return (v + 255) / 256 * 256; }
inline long long conv_chunk_bytes() {
  static const long long v = [] {
    const char* e = std::getenv("PAL_CONV_CHUNK_MB");     // tuning switch
    return (long long)(e ? std::max(1, atoi(e)) : 1 << 20) << 20;   // default: no cap (measured: larger launches win)
  }();
  return v;
}

template <typename T> struct GenericLayout {
  BluePlan p;
  size_t tables;      // bytes of chirp + tw1 + tw2 + twM + bhat
  size_t per_tr;      // conv buffer + two corr rows, per (packed) transform in flight
  size_t per_row;     // one packed spectrum row (two channels)
  GenericLayout(int n) : p(make_blue_plan(n)) {
    // (+ 8 KB: the second-generation float32 engine keeps full-circle stage twiddles, 512 entries at most per axis)
    tables = al(sizeof(cpx<T>) * size_t(p.n)) + al(sizeof(cpx<T>) * (p.M1 / 2 + 1)) +
             al(sizeof(cpx<T>) * (p.M2 / 2 + 1)) + 2 * al(sizeof(cpx<T>) * size_t(p.M)) + 8192;
    per_tr = al(sizeof(cpx<T>) * size_t(p.M)) + 2 * al(sizeof(T) * size_t(p.n));   // a packed inverse yields two rows
    per_row = al(sizeof(cpx<T>) * size_t(p.n + 2));      // a packed spectrum row, or the Hermitian halves of its two channels
  }
};

struct GenericCall {
  const float* sig;
  long long B;
  int Mics, ld, n1, n2;
  const int* pairs;
  int P;
  PickParams pp;
  float eps;
  int* k_idx;
  int* k_count;
  float* peak;
  float* gmax;
  unsigned* flags;
  float* corr_out;
  cudaStream_t stream;
  int sms;
  float* scales;     // [B * Mics][2] per-row power-of-two normalisation (filled by the first sweep, reused by the float64 one)
  const double* sig64 = nullptr;   // float64 rows (pal_gcc_phat_tdoa_f64): the float64 sweep then reads these instead of `sig`
  float* hq = nullptr;             // [B * Mics][2] per-channel bounds of the float32 sweep (pal_winpick.cuh: whiten_unpack_body)
};

// Opt a kernel into `bytes` of dynamic shared memory.  The attribute is per kernel and process-wide, and the first-generation
// engine sizes its tiles at run time: the limit therefore only ever GROWS (a smaller plan on another host thread must not
// shrink it between this call and the launch that relies on it).
inline void grow_smem(const void* kern, size_t bytes) {
  static std::mutex mu;
  static std::map<std::pair<const void*, int>, size_t> cur;
  int dev = 0;
  cudaGetDevice(&dev);
  std::lock_guard<std::mutex> g(mu);
  size_t& c = cur[{kern, dev}];
  if (bytes > c) {
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    c = bytes;
  }
}
inline void count_launch(int k = 1) {
  if (g_launch_counter) g_launch_counter->fetch_add((unsigned long long)(long long)k);
}
inline bool use_fast_pick() {      // PAL_FAST_PICK=0: full find_peaks emulation on every float32 row (A/B measurements)
  static const bool on = [] {
    const char* e = std::getenv("PAL_FAST_PICK");
    return e ? (e[0] != '0') : true;
  }();
  return on;
}
// PAL_FFT2=0 keeps every float32 sweep on the first-generation (power-of-two, run-time-sized) engine: A/B measurements
inline bool use_fft2() {
  static const bool on = [] {
    const char* e = std::getenv("PAL_FFT2");
    return e ? (e[0] != '0') : true;
  }();
  return on;
}

template <typename T> struct BlueBuffers {
  cpx<T>*chirp, *tw1, *tw2, *twM, *bhat;
  BlueTables<T> tb() const { return BlueTables<T>{chirp, tw1, tw2, twM, bhat}; }
};

template <typename T> inline size_t row_smem(const BluePlan& p) {
  const int tr = std::min(p.M1, ColTile<T>::TR);
  return fft_tile_smem(sizeof(T), p.M2, tr + 1);
}
template <typename T> inline int row_units(const BluePlan& p) { return p.M1 / std::min(p.M1, ColTile<T>::TR); }
template <typename T> inline size_t col_smem(const BluePlan& p) {
  const int tc = std::min(p.M2, ColTile<T>::TC);
  return fft_tile_smem(sizeof(T), p.M1, tc);
}

// Row pass of `nt` transforms.  Short rows (M2 <= 128, float32) use tiles of 32 interleaved rows instead of 16: a
// warp then covers 32 different rows at one position, which is free of shared-memory bank conflicts (16-row tiles
// put two butterflies of a warp on overlapping banks), and the tile is still only 34 KB.
#ifndef PAL_WIDE_ROWS
#define PAL_WIDE_ROWS 1
#endif
constexpr int kWideRows = 32;
template <typename T> inline bool wide_rows(const BluePlan& p) {
  return PAL_WIDE_ROWS && sizeof(T) == 4 && p.M2 <= 128 && p.M1 >= kWideRows;
}
// `n_tr_dev` != nullptr: the transform count is read on the device (nt then only sizes the grid)
template <typename T, bool CONV, bool CONJ>
inline void launch_rowpass(const BluePlan& p, const BlueTables<T>& tb, long long nt, cpx<T>* buf, cudaStream_t s, long long max_blocks,
                           const int* n_tr_dev = nullptr) {
  const long long n_arg = n_tr_dev ? 1 : nt;
  if (wide_rows<T>(p)) {
    const size_t sm = fft_tile_smem(sizeof(T), p.M2, kWideRows + 1);
    auto kern = k_rowpass<T, CONV, CONJ, kWideRows>;
    grow_smem(reinterpret_cast<const void*>(kern), sm);
    kern<<<(unsigned)std::min<long long>(nt * (p.M1 / kWideRows), max_blocks), kGT, sm, s>>>(p, tb, n_arg, n_tr_dev, buf);
  } else {
    const size_t sm = row_smem<T>(p);
    auto kern = k_rowpass<T, CONV, CONJ>;
    grow_smem(reinterpret_cast<const void*>(kern), sm);
    kern<<<(unsigned)std::min<long long>(nt * row_units<T>(p), max_blocks), kGT, sm, s>>>(p, tb, n_arg, n_tr_dev, buf);
  }
}

// carve the tables of a plan out of `base` (no launches)
template <typename T> void carve_plan(const BluePlan& p, char*& base, BlueBuffers<T>& bb) {
  bb.chirp = reinterpret_cast<cpx<T>*>(base); base += al(sizeof(cpx<T>) * size_t(p.n));
  bb.tw1 = reinterpret_cast<cpx<T>*>(base);   base += al(sizeof(cpx<T>) * (p.M1 / 2 + 1));
  bb.tw2 = reinterpret_cast<cpx<T>*>(base);   base += al(sizeof(cpx<T>) * (p.M2 / 2 + 1));
  bb.twM = reinterpret_cast<cpx<T>*>(base);   base += al(sizeof(cpx<T>) * size_t(p.M));
  bb.bhat = reinterpret_cast<cpx<T>*>(base);  base += al(sizeof(cpx<T>) * size_t(p.M));
}
// fill them and build the chirp spectrum
template <typename T> cudaError_t fill_plan(const BluePlan& p, const BlueBuffers<T>& bb, cudaStream_t s, int sms) {
  k_blue_init<T><<<std::min(4 * sms, (p.M + kGT - 1) / kGT), kGT, 0, s>>>(p, bb.chirp, bb.tw1, bb.tw2, bb.twM);
  const size_t cs = col_smem<T>(p), rs = row_smem<T>(p);
  grow_smem(reinterpret_cast<const void*>(k_colpass_fwd<T, LoadBhat<T>>), cs);
  const int tiles = p.M2 / std::min(p.M2, ColTile<T>::TC);
  k_colpass_fwd<T, LoadBhat<T>><<<std::min(tiles, 8 * sms), kGT, cs, s>>>(p, bb.tb(), LoadBhat<T>{p, bb.chirp}, 1, nullptr,
                                                                          bb.bhat);
  launch_rowpass<T, false, false>(p, bb.tb(), 1, bb.bhat, s, 8LL * sms);
  count_launch(3);
  return cudaGetLastError();
}
// opt every transform kernel of precision T into the shared memory this plan needs
template <typename T> void plan_kernel_attributes(const BluePlan& p) {
  const size_t cs = col_smem<T>(p), rs = row_smem<T>(p);
  grow_smem(reinterpret_cast<const void*>(k_colpass_fwd<T, LoadSignal<T>>), cs);
  grow_smem(reinterpret_cast<const void*>(k_colpass_fwd<T, LoadSignal2<T>>), cs);
  grow_smem(reinterpret_cast<const void*>(k_colpass_fwd<T, LoadPhat2<T>>), cs);
  grow_smem(reinterpret_cast<const void*>(k_colpass_inv<T, StoreCorr2<T>>), cs);
  grow_smem(reinterpret_cast<const void*>(k_colpass_inv<T, StoreSpectrum<T>>), cs);
  grow_smem(reinterpret_cast<const void*>(k_colpass_inv<T, StoreCorr<T>>), cs);
}
// carve + fill (the one-shot form every non-cached caller uses)
template <typename T> cudaError_t setup_plan(const BluePlan& p, char*& base, BlueBuffers<T>& bb, cudaStream_t s, int sms) {
  carve_plan<T>(p, base, bb);
  plan_kernel_attributes<T>(p);
  return fill_plan<T>(p, bb, s, sms);
}

// ---------------------------------------------------------------- second-generation float32 engine (pal_fft2.cuh)
namespace f2h {
using namespace pal::fft2;
constexpr int kNT2 = 256;
#ifndef PAL_FFT2_MINBLOCKS
#define PAL_FFT2_MINBLOCKS 4
#endif

template <class P, class Loader>
__global__ void __launch_bounds__(kNT2, PAL_FFT2_MINBLOCKS) k2_colpass_fwd(Tables tb, Loader ld, long long n_tr, cpxf* buf) {
  extern __shared__ __align__(128) char smem[];
  colpass_fwd_body<P, kNT2, Loader>(tb, ld, n_tr, buf, smem);
}
template <class P, int MODE>
__global__ void __launch_bounds__(kNT2, PAL_FFT2_MINBLOCKS) k2_rowpass(Tables tb, long long n_tr, cpxf* buf, cpxf* bhat_out) {
  extern __shared__ __align__(128) char smem[];
  rowpass_body<P, kNT2, MODE>(tb, n_tr, buf, bhat_out, smem);
}
template <class P, class Storer>
__global__ void __launch_bounds__(kNT2, PAL_FFT2_MINBLOCKS) k2_colpass_inv(Tables tb, Storer st, long long n_tr, const cpxf* buf) {
  extern __shared__ __align__(128) char smem[];
  colpass_inv_body<P, kNT2, Storer>(tb, st, n_tr, buf, smem);
}
#ifndef PAL_SMEM_NT
#define PAL_SMEM_NT 512
#endif
constexpr int kNTS = PAL_SMEM_NT;   // threads of the single-CTA convolution kernel (one block per SM: 129 KB of shared memory)
template <class P, int MODE, class Loader, class Storer>
__global__ void __launch_bounds__(kNTS, 1) k2_conv_smem(Tables tb, const cpxf* bhat_s, cpxf* bhat_out, Loader ld, Storer st,
                                                        long long n_tr) {
  extern __shared__ __align__(128) char smem[];
  conv_smem_body<P, kNTS, MODE, Loader, Storer>(tb, bhat_s, bhat_out, ld, st, n_tr, smem);
}
// PAL_SMEM_CONV=0: every convolution through the three-kernel engine (A/B measurements)
inline bool use_smem_conv() {
  static const bool on = [] {
    const char* e = std::getenv("PAL_SMEM_CONV");
    return e ? (e[0] != '0') : true;
  }();
  return on;
}
inline bool plan_fits_smem(int plan) {
  bool fits = false;
  with_plan(plan, [&](auto pl) { fits = SmemConv<decltype(pl)>::fits; });
  return fits && use_smem_conv();
}
template <auto kern> inline void opt_in_smem_once(size_t smem) {
  static std::atomic<unsigned> done{0};
  int dev = 0;
  cudaGetDevice(&dev);
  const unsigned bit = 1u << (dev & 31);
  if (!(done.load(std::memory_order_acquire) & bit)) {
    grow_smem(reinterpret_cast<const void*>(kern), smem);
    done.fetch_or(bit, std::memory_order_release);
  }
}

__global__ void __launch_bounds__(256) k2_init_tables(int n, int M1, int M2, cpxf* chirp, cpxf* tw1, cpxf* tw2, cpxf* twf) {
  init_tables_body(n, M1, M2, chirp, tw1, tw2, twf);
}

// Resident blocks per SM of a kernel at its (compile-time) shared-memory size; the opt-in above 48 KB and the
// occupancy query run once per kernel and device, not once per call (and never shrink an attribute another caller set).
// (the kernel is a template ARGUMENT: kernels of equal signature share one function-pointer type, and a cache keyed
// by that type would be shared between them)
template <auto kern> inline int resident_blocks(size_t smem) {
  static std::atomic<int> cache[32];
  int dev = 0;
  cudaGetDevice(&dev);
  dev &= 31;
  int v = cache[dev].load(std::memory_order_acquire);
  if (v > 0) return v;
  grow_smem(reinterpret_cast<const void*>(kern), smem);
  int occ = 1;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, kNT2, smem) != cudaSuccess || occ < 1) occ = 1;
  cache[dev].store(occ, std::memory_order_release);
  return occ;
}
template <auto kern> inline unsigned grid_for(size_t smem, long long units, int sms) {
  return (unsigned)std::max<long long>(1, std::min<long long>(units, (long long)sms * resident_blocks<kern>(smem)));
}

struct Buffers {
  int plan = -1;
  BluePlan p{};          // n and the convolution length (what the loaders / storers look at)
  cpxf *chirp = nullptr, *tw1 = nullptr, *tw2 = nullptr, *twf = nullptr, *bhat = nullptr;

  Tables tb() const { return Tables{chirp, tw1, tw2, twf, bhat}; }
};
inline size_t table_bytes(int n, int plan) {
  const PlanDims d = plan_dims(plan);
  return al(sizeof(cpxf) * size_t(n)) + al(sizeof(cpxf) * d.M1) + al(sizeof(cpxf) * d.M2) + 2 * al(sizeof(cpxf) * size_t(d.M1) * d.M2);
}

inline void carve(int n, int plan, char*& base, Buffers& b) {
  const PlanDims d = plan_dims(plan);
  const size_t M = size_t(d.M1) * d.M2;
  b.plan = plan;
  b.p = BluePlan{n, int(M), d.M1, d.M2, 0, 0};
  b.chirp = reinterpret_cast<cpxf*>(base); base += al(sizeof(cpxf) * size_t(n));
  b.tw1 = reinterpret_cast<cpxf*>(base);   base += al(sizeof(cpxf) * d.M1);
  b.tw2 = reinterpret_cast<cpxf*>(base);   base += al(sizeof(cpxf) * d.M2);
  b.twf = reinterpret_cast<cpxf*>(base);   base += al(sizeof(cpxf) * M);
  b.bhat = reinterpret_cast<cpxf*>(base);  base += al(sizeof(cpxf) * M);
}
// the single-CTA kernel needs only ONE chirp-spectrum table: plans that fit shared memory keep it in `bhat`'s slot
inline bool smem_plan(const Buffers& b) { return b.plan >= 0 && plan_fits_smem(b.plan); }
// chirp, twiddles and the chirp spectrum; `scratch` holds one convolution (M complex)
inline cudaError_t fill(const Buffers& b, cpxf* scratch, cudaStream_t s, int sms) {
  const PlanDims d = plan_dims(b.plan);
  k2_init_tables<<<std::min(4 * sms, (b.p.M + 255) / 256), 256, 0, s>>>(b.p.n, d.M1, d.M2, b.chirp, b.tw1, b.tw2, b.twf);
  with_plan(b.plan, [&](auto pl) {
    using P = decltype(pl);
    if constexpr (SmemConv<P>::fits) {
      if (use_smem_conv()) {       // the chirp spectrum in the single-CTA kernel's order, built by that kernel
        using K = StoreRaw<float>;
        opt_in_smem_once<k2_conv_smem<P, 2, LoadBhat<float>, K>>(SmemConv<P>::smem);
        k2_conv_smem<P, 2, LoadBhat<float>, K><<<1, kNTS, SmemConv<P>::smem, s>>>(b.tb(), nullptr, b.bhat, LoadBhat<float>{b.p, b.chirp},
                                                                                 K{scratch, (long long)P::M}, 1);
        return;
      }
    }
    k2_colpass_fwd<P, LoadBhat<float>><<<grid_for<k2_colpass_fwd<P, LoadBhat<float>>>(P::col_smem, P::M2 / P::TC, sms), kNT2, P::col_smem, s>>>(
        b.tb(), LoadBhat<float>{b.p, b.chirp}, 1, scratch);
    k2_rowpass<P, 2><<<grid_for<k2_rowpass<P, 2>>(P::row_smem, P::M1 / P::TR, sms), kNT2, P::row_smem, s>>>(b.tb(), 1, scratch, b.bhat);
  });
  count_launch(3);
  return cudaGetLastError();
}
// one batch of `nt` convolutions: loader -> (x chirp spectrum, conjugated when CONJ) -> storer
template <bool CONJ, class Loader, class Storer>
inline void conv(const Buffers& b, const Loader& ld, const Storer& st, long long nt, cpxf* buf, cudaStream_t s, int sms) {
  with_plan(b.plan, [&](auto pl) {
    using P = decltype(pl);
    if constexpr (SmemConv<P>::fits) {
      if (use_smem_conv()) {       // whole convolution in one CTA's shared memory (b.bhat is in that kernel's order)
        opt_in_smem_once<k2_conv_smem<P, CONJ ? 1 : 0, Loader, Storer>>(SmemConv<P>::smem);
        k2_conv_smem<P, CONJ ? 1 : 0, Loader, Storer><<<(unsigned)std::max<long long>(1, std::min<long long>(nt, sms)), kNTS,
                                                        SmemConv<P>::smem, s>>>(b.tb(), b.bhat, nullptr, ld, st, nt);
        return;
      }
    }
    k2_colpass_fwd<P, Loader><<<grid_for<k2_colpass_fwd<P, Loader>>(P::col_smem, nt * (P::M2 / P::TC), sms), kNT2, P::col_smem, s>>>(
        b.tb(), ld, nt, buf);
    k2_rowpass<P, CONJ ? 1 : 0><<<grid_for<k2_rowpass<P, CONJ ? 1 : 0>>(P::row_smem, nt * (P::M1 / P::TR), sms), kNT2, P::row_smem, s>>>(
        b.tb(), nt, buf, nullptr);
    k2_colpass_inv<P, Storer><<<grid_for<k2_colpass_inv<P, Storer>>(P::col_smem, nt * (P::M2 / P::TC), sms), kNT2, P::col_smem, s>>>(
        b.tb(), st, nt, buf);
  });
  count_launch(3);
}
}  // namespace f2h

#ifndef PAL_PICK_BLOCKS
#define PAL_PICK_BLOCKS 4
#endif
inline int pick_grid(int sms) { return std::max(1, std::min(PAL_PICK_BLOCKS * sms, 1024)); }   // 60 registers x 256 threads: four blocks per SM
inline size_t pkmap_bytes(int n, int sms) { return al(size_t((n + 15) / 16 * 16) * pick_grid(sms)); }

// smallest / comfortable workspace of one sweep in precision T
template <typename T> size_t generic_min_bytes(int n, int rows_per_unit, int sms) {
  GenericLayout<T> L(n);
  return L.tables + pkmap_bytes(n, sms) + L.per_tr + size_t(rows_per_unit) * L.per_row + 1024;
}
template <typename T> size_t generic_full_bytes(int n, long long B, int Mics, int P, int sms) {
  GenericLayout<T> L(n);
  const long long tr = std::min<long long>(2048, std::max<long long>(B * P, B * Mics));
  return L.tables + pkmap_bytes(n, sms) + size_t(tr) * L.per_tr + size_t(B) * Mics * L.per_row + 1024;
}

// One sweep of the whole batch in precision T.  items: all B*P (list == nullptr) or the
// `n_list` flagged items of `list` (host-known count), whose channel rows are in `rows_scratch`.
// Two real sequences share every complex transform (pal_bluestein.cuh): the channels of a frame are
// transformed in pairs (CP = ceil(Mics/2) packed spectrum rows per frame; in list mode the two channels of
// an item form one row), and every inverse transform yields the correlation rows of two items.
// `count_dev` (list mode only): the list length lives on the device.  n_list is then an upper bound, the sweep runs in
// rounds whose sizes are derived on the device (k_chunk_counts -> dev_items / dev_packed, kMaxDevRounds ints each) and
// every kernel reads its own count: no host round trip.  Returns cudaErrorNotSupported when the workspace would need
// more than kMaxDevRounds rounds (the caller then reads the count back and uses the host-counted form).
template <typename T>
cudaError_t run_generic(const GenericCall& c, char* ws, size_t ws_bytes, const int* list, int n_list, int* rows_scratch,
                        unsigned extra_flag, unsigned keep_mask, const int* count_dev = nullptr, int* dev_items = nullptr,
                        int* dev_packed = nullptr) {
  const int n = c.n1 + c.n2 - 1;
  GenericLayout<T> L(n);
  const BluePlan p = L.p;
  const int CP = (c.Mics + 1) / 2;
  if (ws_bytes < generic_min_bytes<T>(n, list ? 2 : c.Mics, c.sms)) return cudaErrorMemoryAllocation;
  char* base = ws;
  BlueBuffers<T> bb;
  // float32 sweeps run on the second-generation engine whenever one of its plans holds 2n - 1 points
  f2h::Buffers b2;
  const int plan2 = (std::is_same<T, float>::value && use_fft2()) ? fft2::choose_plan(n) : -1;
  cudaError_t e = cudaSuccess;
  if (plan2 >= 0) {
    f2h::carve(n, plan2, base, b2);
    bb.chirp = reinterpret_cast<cpx<T>*>(b2.chirp);
    base = ws + L.tables;        // same budget as the first-generation tables (never smaller)
  } else {
    e = setup_plan<T>(p, base, bb, c.stream, c.sms);
    if (e != cudaSuccess) return e;
  }
  const int grid_pick = pick_grid(c.sms);
  unsigned char* pkmap = reinterpret_cast<unsigned char*>(base);
  base += pkmap_bytes(n, c.sms);
  size_t rem = ws_bytes - size_t(base - ws);
  // packed transforms in flight: a quarter of what is left (at most 2048), the rest holds spectrum rows
  const long long total_items = list ? n_list : c.B * c.P;
  const long long min_rows = list ? 1 : CP;
  long long tr_cap = std::max<long long>(1, std::min<long long>(2048, (long long)((rem / 4) / L.per_tr)));
  // optional cap on the convolution buffers of one chunk (PAL_CONV_CHUNK_MB; L2-sized chunks measured slower)
  tr_cap = std::max<long long>(1, std::min<long long>(tr_cap, conv_chunk_bytes() / (long long)(sizeof(cpx<T>) * size_t(p.M))));
  tr_cap = std::min<long long>(tr_cap, std::max<long long>((total_items + 1) / 2, list ? (long long)n_list : c.B * CP));
  while (tr_cap > 1 && rem < size_t(tr_cap) * L.per_tr + size_t(min_rows) * L.per_row) tr_cap /= 2;
  // reduced pick (pal_winpick.cuh): one peak, no correlation rows wanted, near-tie audit on, first sweep only.  Its forward
  // sub-chunks must start on frame boundaries (the channels of a frame are whitened together).
  const int c0 = c.n2 - 1;
  const WinGeom wg = make_win_geom(n, c0, c.pp.win_half, c.pp.dist, c.eps);
  // partial row maxima per row: one per column tile, or a single one when the whole convolution runs in one CTA
  const int win_tiles = plan2 < 0 ? 0 : (f2h::plan_fits_smem(plan2) ? 1 : fft2::plan_dims(plan2).M2 / (fft2::plan_dims(plan2).M1 <= 192 ? 32 : 16));
  const bool fast_pick = plan2 >= 0 && !list && c.pp.num_peaks == 1 && !c.corr_out && c.eps > 0.f && use_fast_pick() &&
                         size_t(wg.wstride + win_tiles) * sizeof(float) <= al(sizeof(T) * size_t(n)) && tr_cap >= CP;
  // per-channel whitening pays when a channel is used by several pairs (cfg5: 28 pairs / 8 channels, cfg4: 2016 / 64);
  // with few pairs per channel (cfg2: 6 / 4) the extra pass over the spectra costs more than the leaner pair loader saves
  const bool whiten = fast_pick && c.P >= 2 * c.Mics;
  if (whiten) tr_cap -= tr_cap % CP;
  // near-tie audit on: the rounding noise of the float32 forward transforms is bounded per channel (and, on whitened
  // spectra, the neglected factor g); [resident frames * Mics][2], see whiten_unpack_body
  float* hq_res = nullptr;
  if constexpr (std::is_same<T, float>::value) {
    if (!list && c.eps > 0.f) hq_res = c.hq;
  }
  if (whiten && !hq_res) return cudaErrorInvalidValue;
  cpx<T>* conv = reinterpret_cast<cpx<T>*>(base);
  base += tr_cap * al(sizeof(cpx<T>) * size_t(p.M));
  T* corr = reinterpret_cast<T*>(base);
  base += 2 * tr_cap * al(sizeof(T) * size_t(n));
  rem = ws_bytes - size_t(base - ws);
  long long row_cap = (long long)(rem / L.per_row);
  if (row_cap < min_rows) return cudaErrorMemoryAllocation;
  if (count_dev) {
    // one round = one pass of every kernel: forward transforms (one per item) and inverse transforms share the chunk.
    // The rounds must cover the worst case (every item flagged) although a fraction of a percent is the rule, so the
    // chunk is made as large as the workspace allows: rounds that find nothing cost a few microseconds each.
    const long long fit = (long long)((ws_bytes - size_t(reinterpret_cast<char*>(conv) - ws)) / (L.per_tr + L.per_row));
    const long long chunk = std::max<long long>(1, std::min<long long>(fit, 32768));
    const long long rounds = (n_list + chunk - 1) / chunk;
    if (rounds > kMaxDevRounds) return cudaErrorNotSupported;
    tr_cap = row_cap = chunk;
    base = reinterpret_cast<char*>(conv) + chunk * al(sizeof(cpx<T>) * size_t(p.M));
    corr = reinterpret_cast<T*>(base);
    base += 2 * chunk * al(sizeof(T) * size_t(n));
    k_chunk_counts<<<1, 128, 0, c.stream>>>(count_dev, int(chunk), int(rounds), dev_items, dev_packed);
    count_launch();
  }
  cpx<T>* spec = reinterpret_cast<cpx<T>*>(base);
  // note: conv rows / corr rows / spectrum rows are addressed densely (t * M, t * n), the al()
  // padding above only makes the regions start aligned
  const size_t cs = col_smem<T>(p), rs = row_smem<T>(p);
  const int tiles = p.M2 / std::min(p.M2, ColTile<T>::TC);
  const BlueTables<T> tb = bb.tb();
  const BluePlan pl = plan2 >= 0 ? b2.p : p;        // what loaders / storers see: n and the convolution length
  if (plan2 >= 0) {
    if constexpr (std::is_same<T, float>::value) {
      e = f2h::fill(b2, reinterpret_cast<cpxf*>(conv), c.stream, c.sms);
      if (e != cudaSuccess) return e;
    }
  }
  if (!list) {
    if (c.sig64)
      k_row_scales<double><<<(unsigned)std::min<long long>(c.B * c.Mics, 16LL * c.sms), kGT, 0, c.stream>>>(c.sig64, c.B * c.Mics, c.ld,
                                                                                                          c.n1, c.n2, c.scales);
    else
      k_row_scales<float><<<(unsigned)std::min<long long>(c.B * c.Mics, 16LL * c.sms), kGT, 0, c.stream>>>(c.sig, c.B * c.Mics, c.ld, c.n1,
                                                                                                         c.n2, c.scales);
    count_launch();
  }

  // packed forward transforms g0 .. g0+ntr-1 (global packed index; list mode: flagged-item index) -> spec_out rows 0..
  int dev_round = 0;      // device-counted list mode: index of the round being issued
  auto forward = [&](long long g0, long long ntr, const int* row_list, cpx<T>* spec_out) {
    for (long long r0 = 0; r0 < ntr; r0 += tr_cap) {
      const long long nt = std::min(tr_cap, ntr - r0);
      if (count_dev) {       // every kernel of the round reads its transform count on the device
        LoadSignal2<T> ldd{pl, bb.chirp, c.sig, c.ld, c.Mics, CP, c.n1, c.n2, row_list, g0 + r0, c.scales};
        k_colpass_fwd<T, LoadSignal2<T>><<<(unsigned)std::min<long long>(nt * tiles, 16LL * c.sms), kGT, cs, c.stream>>>(
            p, tb, ldd, 1, dev_items + dev_round, conv);
        launch_rowpass<T, true, false>(p, tb, nt, conv, c.stream, 16LL * c.sms, dev_items + dev_round);
        StoreSpectrum<T> std_{p, bb.chirp, spec_out + r0 * n};
        k_colpass_inv<T, StoreSpectrum<T>><<<(unsigned)std::min<long long>(nt * tiles, 16LL * c.sms), kGT, cs, c.stream>>>(
            p, tb, std_, 1, dev_items + dev_round, conv);
        count_launch(3);
        continue;
      }
      if constexpr (std::is_same<T, double>::value) {
        if (c.sig64) {       // float64 rows straight into the float64 transforms
          LoadSignal2<T, double> ld64{pl, bb.chirp, c.sig64, c.ld, c.Mics, CP, c.n1, c.n2, row_list, g0 + r0, c.scales};
          grow_smem(reinterpret_cast<const void*>(k_colpass_fwd<T, LoadSignal2<T, double>>), cs);
          k_colpass_fwd<T, LoadSignal2<T, double>><<<(unsigned)std::min<long long>(nt * tiles, 16LL * c.sms), kGT, cs, c.stream>>>(
              p, tb, ld64, nt, nullptr, conv);
          launch_rowpass<T, true, false>(p, tb, nt, conv, c.stream, 16LL * c.sms);
          StoreSpectrum<T> st64{p, bb.chirp, spec_out + r0 * n};
          k_colpass_inv<T, StoreSpectrum<T>><<<(unsigned)std::min<long long>(nt * tiles, 16LL * c.sms), kGT, cs, c.stream>>>(
              p, tb, st64, nt, nullptr, conv);
          count_launch(3);
          continue;
        }
      }
      LoadSignal2<T> ld{pl, bb.chirp, c.sig, c.ld, c.Mics, CP, c.n1, c.n2, row_list, g0 + r0, c.scales};
      if constexpr (std::is_same<T, float>::value) {
        if (whiten) {
          // packed spectra of this sub-chunk into the (still unused) correlation-row region, then every channel
          // unpacked, whitened and halved into the resident spectrum rows (pal_winpick.cuh: whiten_unpack_body)
          cpxf* ztmp = reinterpret_cast<cpxf*>(corr);
          f2h::conv<false>(b2, ld, StoreSpectrum<T>{pl, bb.chirp, ztmp}, nt, reinterpret_cast<cpxf*>(conv), c.stream, c.sms);
          const long long frame_first = (g0 + r0) / CP;           // sub-chunks start on frame boundaries (see below)
          k_whiten_unpack<<<(unsigned)std::min<long long>(nt, 16LL * c.sms), kGT, 0, c.stream>>>(
              ztmp, n, nt, c.Mics, CP, c.scales, frame_first * c.Mics, (frame_first - g0 / CP) * c.Mics,
              reinterpret_cast<cpxf*>(spec_out), hq_res);
          count_launch();
          continue;
        }
        if (plan2 >= 0) {
          f2h::conv<false>(b2, ld, StoreSpectrum<T>{pl, bb.chirp, spec_out + r0 * n}, nt, reinterpret_cast<cpxf*>(conv), c.stream, c.sms);
          continue;
        }
      }
      k_colpass_fwd<T, LoadSignal2<T>><<<(unsigned)std::min<long long>(nt * tiles, 16LL * c.sms), kGT, cs, c.stream>>>(
          p, tb, ld, nt, nullptr, conv);
      launch_rowpass<T, true, false>(p, tb, nt, conv, c.stream, 16LL * c.sms);
      StoreSpectrum<T> st{p, bb.chirp, spec_out + r0 * n};
      k_colpass_inv<T, StoreSpectrum<T>><<<(unsigned)std::min<long long>(nt * tiles, 16LL * c.sms), kGT, cs, c.stream>>>(
          p, tb, st, nt, nullptr, conv);
      count_launch(3);
    }
  };
  // items [0, nitems) of the resident set (frame-major, or flagged-list order when ilist != nullptr);
  // out0 = global item id of resident item 0 (frame-major mode)
  auto inverse = [&](long long nitems, const cpx<T>* spec_in, long long out0, const int* ilist, long long frame0, long long list0) {
    for (long long i0 = 0; i0 < nitems; i0 += 2 * tr_cap) {
      const long long ni = std::min(2 * tr_cap, nitems - i0);
      const long long nt = (ni + 1) / 2;
      LoadPhat2<T> ld{pl, bb.chirp, spec_in, c.pairs, c.Mics, CP, c.P, i0, nitems, ilist != nullptr, c.scales, frame0, rows_scratch, list0};
      if (count_dev) ld.n_items_dev = dev_items + dev_round;
      StoreCorr2<T> st{pl, bb.chirp, corr, ni, ld};
      bool done2 = false;
      if (count_dev) {
        k_colpass_fwd<T, LoadPhat2<T>><<<(unsigned)std::min<long long>(nt * tiles, 16LL * c.sms), kGT, cs, c.stream>>>(
            p, tb, ld, 1, dev_packed + dev_round, conv);
        launch_rowpass<T, true, true>(p, tb, nt, conv, c.stream, 16LL * c.sms, dev_packed + dev_round);
        k_colpass_inv<T, StoreCorr2<T>><<<(unsigned)std::min<long long>(nt * tiles, 16LL * c.sms), kGT, cs, c.stream>>>(
            p, tb, st, 1, dev_packed + dev_round, conv);
        k_pick_rows<T><<<(unsigned)std::min<long long>(ni, grid_pick), kGT, sizeof(RowPickSmem), c.stream>>>(
            corr, n, c0, ni, dev_items + dev_round, ilist + i0, c.pp.win_half, c.pp.dist, c.pp.method, c.pp.mult, c.pp.num_peaks,
            c.eps, pkmap, c.k_idx, c.k_count, c.peak, c.gmax, c.flags, extra_flag, keep_mask, nullptr);
        count_launch(4);
        continue;
      }
      if constexpr (std::is_same<T, float>::value) {
        if (fast_pick && !whiten) {
          // only the window (+ margin) and per-tile row maxima leave the inverse column pass; one warp per row picks
          float* win = reinterpret_cast<float*>(corr);
          float* pmax = win + size_t(2 * tr_cap) * wg.wstride;
          StoreWin2 sw{pl, b2.chirp, win, pmax, ni, ld, wg, win_tiles};
          f2h::conv<true>(b2, ld, sw, nt, reinterpret_cast<cpxf*>(conv), c.stream, c.sms);
          const long long o = out0 + i0;
          k_win_pick<<<(unsigned)std::min<long long>((ni + kGT / 32 - 1) / (kGT / 32), 8LL * c.sms), kGT, 0, c.stream>>>(
              win, pmax, win_tiles, ni, wg, o, c.k_idx, c.k_count, c.peak, c.gmax, c.flags, extra_flag,
              WhitenRef{hq_res, c.pairs, c.Mics, c.P, i0, false});
          count_launch(1);
          continue;
        }
        if (fast_pick) {
          // whitened half spectra in, only the window (+ margin) and per-tile row maxima out; one warp per row picks
          float* win = reinterpret_cast<float*>(corr);
          float* pmax = win + size_t(2 * tr_cap) * wg.wstride;
          const LoadPhatU lu{pl, b2.chirp, reinterpret_cast<const cpxf*>(spec_in), c.pairs, c.Mics, c.P, n / 2 + 1, i0, nitems,
                             c.scales, frame0};
          StoreWinU sw{pl, b2.chirp, win, pmax, ni, lu, wg, win_tiles};
          f2h::conv<true>(b2, lu, sw, nt, reinterpret_cast<cpxf*>(conv), c.stream, c.sms);
          const long long o = out0 + i0;
          k_win_pick<<<(unsigned)std::min<long long>((ni + kGT / 32 - 1) / (kGT / 32), 8LL * c.sms), kGT, 0, c.stream>>>(
              win, pmax, win_tiles, ni, wg, o, c.k_idx, c.k_count, c.peak, c.gmax, c.flags, extra_flag,
              WhitenRef{hq_res, c.pairs, c.Mics, c.P, i0});
          count_launch(1);
          continue;
        }
        if (plan2 >= 0) {
          f2h::conv<true>(b2, ld, st, nt, reinterpret_cast<cpxf*>(conv), c.stream, c.sms);
          count_launch(-2);      // (the four launches of this chunk are counted below)
          done2 = true;
        }
      }
      if (!done2) {
        k_colpass_fwd<T, LoadPhat2<T>><<<(unsigned)std::min<long long>(nt * tiles, 16LL * c.sms), kGT, cs, c.stream>>>(
            p, tb, ld, nt, nullptr, conv);
        launch_rowpass<T, true, true>(p, tb, nt, conv, c.stream, 16LL * c.sms);
        k_colpass_inv<T, StoreCorr2<T>><<<(unsigned)std::min<long long>(nt * tiles, 16LL * c.sms), kGT, cs, c.stream>>>(
            p, tb, st, nt, nullptr, conv);
      }
      const long long o = ilist ? 0 : out0 + i0;
      k_pick_rows<T><<<(unsigned)std::min<long long>(ni, grid_pick), kGT, sizeof(RowPickSmem), c.stream>>>(
          corr, n, c0, ni, nullptr, ilist ? ilist + i0 : nullptr, c.pp.win_half, c.pp.dist, c.pp.method, c.pp.mult,
          c.pp.num_peaks, c.eps, pkmap, c.k_idx + o * c.pp.num_peaks, c.k_count ? c.k_count + o : nullptr, c.peak + o,
          c.gmax + o, c.flags + o, extra_flag, keep_mask, (c.corr_out && !ilist) ? c.corr_out + o * n : nullptr);
      count_launch(4);
    }
  };

  if (!list) {
    const long long fchunk = std::max<long long>(1, std::min<long long>(c.B, row_cap / CP));
    for (long long f0 = 0; f0 < c.B; f0 += fchunk) {
      const long long nf = std::min(fchunk, c.B - f0);
      forward(f0 * CP, nf * CP, nullptr, spec);
      if constexpr (std::is_same<T, float>::value) {
        if (hq_res && !whiten) {      // statistics of the packed spectra (the whitened sweep gathers them while it unpacks)
          k_whiten_unpack<<<(unsigned)std::min<long long>(nf * CP, 16LL * c.sms), kGT, 0, c.stream>>>(
              reinterpret_cast<const cpxf*>(spec), n, nf * CP, c.Mics, CP, c.scales, f0 * c.Mics, 0, nullptr, hq_res);
          count_launch();
        }
      }
      // the resident spectra start at frame f0: item ids handed to the loader are relative to it
      inverse(nf * c.P, spec, f0 * c.P, nullptr, f0, 0);
      if constexpr (std::is_same<T, float>::value) {
        if (hq_res && !fast_pick) {
          k_noise_flags<<<(unsigned)((nf * c.P + 255) / 256), 256, 0, c.stream>>>(WhitenRef{hq_res, c.pairs, c.Mics, c.P, 0, false},
                                                                                nf * c.P, n, c.eps, c.flags + f0 * c.P);
          count_launch();
        }
      }
    }
  } else {
    const long long ichunk = std::max<long long>(1, std::min<long long>(n_list, row_cap));
    for (long long i0 = 0; i0 < n_list; i0 += ichunk) {
      const long long ni = std::min<long long>(ichunk, n_list - i0);
      forward(i0, ni, rows_scratch, spec);
      inverse(ni, spec, 0, list + i0, 0, i0);
      ++dev_round;
    }
  }
  return cudaGetLastError();
}

// ---------------------------------------------------------------- alignment front end (pal_sync.cuh)
constexpr int kSyncThreads = 256;
__global__ void __launch_bounds__(kSyncThreads) k_sync_energy(const double* sig, long long n_scenes, int Mics, long long ld,
                                                              int len, const int* lens, double* energy, int* ref_idx) {
  extern __shared__ __align__(128) char smem[];
  sync_energy_body<kSyncThreads>(sig, n_scenes, Mics, ld, len, lens, energy, ref_idx, smem);
}
template <typename T>
__global__ void __launch_bounds__(kSyncThreads) k_sync_pick(const T* corr, int n, long long n_rows, long long item0, int Mics,
                                                            const int* ref, int len, const int* lens, int* peak_index,
                                                            double* absmax, double* win) {
  __shared__ SyncPickSmem sm;
  sync_pick_body<T, kSyncThreads>(corr, n, n_rows, item0, Mics, ref, len, lens, peak_index, absmax, win,
                                  reinterpret_cast<char*>(&sm));
}
template <typename T>
__global__ void __launch_bounds__(256) k_pad_rows(const T* in, long long n_rows, long long ld_in, int len, const int* lens,
                                                  const int* pad, T* out, long long ld_out) {
  pad_rows_body<T>(in, n_rows, ld_in, len, lens, pad, out, ld_out);
}

struct SyncCall {
  const double* sig;     // [S][Mics][ld]
  long long S;
  int Mics, ld;
  const int* lens;       // optional [S][Mics]
  int* ref_idx;          // out [S]
  int* peak_index;       // out [S][Mics]
  double* absmax;        // out [S][Mics]
  double* win;           // out [S][Mics][5]
  double* energy;        // out [S][Mics] or NULL
  cudaStream_t stream;
  int sms;
};
template <typename T> size_t sync_min_bytes(int ld, int Mics) {
  GenericLayout<T> L(2 * ld - 1);
  return L.tables + L.per_tr + size_t(Mics) * L.per_row + 1024;
}
template <typename T> size_t sync_full_bytes(int ld, long long S, int Mics) {
  GenericLayout<T> L(2 * ld - 1);
  const long long tr = std::min<long long>(2048, std::max<long long>(1, S * Mics));
  return L.tables + size_t(tr) * L.per_tr + size_t(std::max<long long>(1, S)) * Mics * L.per_row + 1024;
}

// energies -> reference channel -> spectra of every channel -> plain cross-correlation with the reference -> arg-max
template <typename T> cudaError_t run_sync_align(const SyncCall& c, char* ws, size_t ws_bytes) {
  const int n = 2 * c.ld - 1;
  GenericLayout<T> L(n);
  const BluePlan p = L.p;
  if (ws_bytes < sync_min_bytes<T>(c.ld, c.Mics)) return cudaErrorMemoryAllocation;
  k_sync_energy<<<(unsigned)std::min<long long>(c.S, 8LL * c.sms), kSyncThreads, sizeof(double) * c.Mics, c.stream>>>(
      c.sig, c.S, c.Mics, c.ld, c.ld, c.lens, c.energy, c.ref_idx);
  count_launch();
  char* base = ws;
  BlueBuffers<T> bb;
  cudaError_t e = setup_plan<T>(p, base, bb, c.stream, c.sms);
  if (e != cudaSuccess) return e;
  const size_t cs = col_smem<T>(p), rs = row_smem<T>(p);
  grow_smem(reinterpret_cast<const void*>(k_colpass_fwd<T, LoadSignalF64<T>>), cs);
  grow_smem(reinterpret_cast<const void*>(k_colpass_fwd<T, LoadCross<T>>), cs);
  size_t rem = ws_bytes - size_t(base - ws);
  long long tr_cap = std::max<long long>(1, std::min<long long>(2048, (long long)((rem / 4) / L.per_tr)));
  tr_cap = std::min<long long>(tr_cap, c.S * c.Mics);
  while (tr_cap > 1 && rem < size_t(tr_cap) * L.per_tr + size_t(c.Mics) * L.per_row) tr_cap /= 2;
  cpx<T>* conv = reinterpret_cast<cpx<T>*>(base);
  base += tr_cap * al(sizeof(cpx<T>) * size_t(p.M));
  T* corr = reinterpret_cast<T*>(base);
  base += tr_cap * al(sizeof(T) * size_t(n));
  rem = ws_bytes - size_t(base - ws);
  const long long row_cap = (long long)(rem / L.per_row);
  if (row_cap < c.Mics) return cudaErrorMemoryAllocation;
  cpx<T>* spec = reinterpret_cast<cpx<T>*>(base);
  const int tiles = p.M2 / std::min(p.M2, ColTile<T>::TC);
  const BlueTables<T> tb = bb.tb();
  const long long schunk = std::max<long long>(1, std::min<long long>(c.S, row_cap / c.Mics));
  for (long long s0 = 0; s0 < c.S; s0 += schunk) {
    const long long ns = std::min(schunk, c.S - s0);
    const long long nrows = ns * c.Mics;
    for (long long r0 = 0; r0 < nrows; r0 += tr_cap) {
      const long long nt = std::min(tr_cap, nrows - r0);
      LoadSignalF64<T> ld{p, bb.chirp, c.sig, c.ld, c.ld, c.lens, s0 * c.Mics + r0};
      k_colpass_fwd<T, LoadSignalF64<T>><<<(unsigned)std::min<long long>(nt * tiles, 16LL * c.sms), kGT, cs, c.stream>>>(
          p, tb, ld, nt, nullptr, conv);
      launch_rowpass<T, true, false>(p, tb, nt, conv, c.stream, 16LL * c.sms);
      StoreSpectrum<T> st{p, bb.chirp, spec + r0 * n};
      k_colpass_inv<T, StoreSpectrum<T>><<<(unsigned)std::min<long long>(nt * tiles, 16LL * c.sms), kGT, cs, c.stream>>>(
          p, tb, st, nt, nullptr, conv);
      count_launch(3);
    }
    for (long long i0 = 0; i0 < nrows; i0 += tr_cap) {
      const long long nt = std::min(tr_cap, nrows - i0);
      LoadCross<T> ld{p, bb.chirp, spec, c.ref_idx, s0, c.Mics, i0};
      k_colpass_fwd<T, LoadCross<T>><<<(unsigned)std::min<long long>(nt * tiles, 16LL * c.sms), kGT, cs, c.stream>>>(
          p, tb, ld, nt, nullptr, conv);
      launch_rowpass<T, true, true>(p, tb, nt, conv, c.stream, 16LL * c.sms);
      StoreCorr<T> st{p, bb.chirp, corr};
      k_colpass_inv<T, StoreCorr<T>><<<(unsigned)std::min<long long>(nt * tiles, 16LL * c.sms), kGT, cs, c.stream>>>(
          p, tb, st, nt, nullptr, conv);
      k_sync_pick<T><<<(unsigned)std::min<long long>(nt, 8LL * c.sms), kSyncThreads, 0, c.stream>>>(
          corr, n, nt, s0 * c.Mics + i0, c.Mics, c.ref_idx, c.ld, c.lens, c.peak_index, c.absmax, c.win);
      count_launch(4);
    }
  }
  return cudaGetLastError();
}

}  // namespace palhost
