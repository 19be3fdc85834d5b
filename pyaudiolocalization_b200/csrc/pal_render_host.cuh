// pal_render_host.cuh -- host-side orchestration of stage 1 (image sources, multipath renderer).
// Included by pal_capi.cu only.
#pragma once
#include "pal_generic_host.cuh"
#include "pal_render.cuh"

#include <cstring>
#include <vector>

namespace palhost {

constexpr int kImgThreads = 128;
constexpr int kXferJ = 16;

__global__ void __launch_bounds__(kImgThreads)
    k_image_sources(ImgParams ip, const double* sources, long long n_scenes, const double* planes, long long plane_stride,
                    const int* plane_mat,
                    const double* mat_abs, const double* mat_freq, const double* mics, long long mic_stride,
                    double* out_pos, int* out_mat, int* out_count, char* scratch, size_t per_block) {
  extern __shared__ __align__(16) char smem[];
  image_sources_body<kImgThreads>(ip, sources, n_scenes, planes, plane_stride, plane_mat, mat_abs, mat_freq, mics, mic_stride, out_pos,
                                  out_mat, out_count, scratch, per_block, smem);
}
__global__ void k_path_table(const double* src, const double* img_pos, const int* img_mat, int n_img, const double* mics,
                             int n_mics, const double* mat_abs, const double* mat_freq, int air_mat, double frequency,
                             double c_sound, double* tau, double* gain) {
  path_table_body(src, img_pos, img_mat, n_img, mics, n_mics, mat_abs, mat_freq, air_mat, frequency, c_sound, tau, gain);
}
__global__ void __launch_bounds__(kGT) k_transfer(const cpxf* X, int N, RenderRows rr, long long row0, long long n_rows,
                                                  double fs, cpxf* G, int* live) {
  extern __shared__ __align__(16) char smem[];
  transfer_body<kGT, kXferJ>(X, N, rr, row0, n_rows, fs, G, live, smem);
}
__global__ void __launch_bounds__(128) k_path_table_batched(const double* sources, const double* img_pos, const int* img_mat,
                                                            const int* img_count, long long n_scenes, int k_max,
                                                            const double* mics, int n_mics, long long mic_stride,
                                                            const double* mat_abs, const double* mat_freq, int air_mat,
                                                            double frequency, double c_sound, int k_stride, double* tau,
                                                            double* gain, int* path_count, double* max_tau) {
  extern __shared__ __align__(16) char smem[];
  path_table_batched_body<128>(sources, img_pos, img_mat, img_count, n_scenes, k_max, mics, n_mics, mic_stride, mat_abs,
                               mat_freq, air_mat, frequency, c_sound, k_stride, tau, gain, path_count, max_tau, smem);
}
__global__ void __launch_bounds__(kGT) k_normalise_compress(float* rows, long long n_rows, int n, float thr, float eps,
                                                            bool compress) {
  extern __shared__ __align__(16) char smem[];
  normalise_compress_body<kGT>(rows, n_rows, n, thr, eps, compress, smem);
}

inline size_t image_scratch_per_block(int n_planes, int k_max) {
  const size_t cmax = size_t(k_max) * n_planes;
  return al(cmax * 3 * 8 * 2 + size_t(k_max + 1) * 3 * 8 + cmax * 4 + 64);
}
inline int image_grid(long long n_scenes, int sms) { return (int)std::min<long long>(n_scenes, 8LL * sms); }

// workspace of a render call over `rows` (scene, mic) rows of transform length 2N:
// tables(2N) + X[2N] + per row in flight: G[N+1] + one convolution buffer
inline size_t render_row_bytes(int N) {
  GenericLayout<float> L(2 * N);
  return al(sizeof(cpxf) * size_t(N + 1)) + al(sizeof(cpxf) * size_t(L.p.M)) + sizeof(int);   // G row, convolution buffer, live flag
}
inline size_t render_plan_bytes(int N);
inline size_t render_fixed_bytes(int N) { return render_plan_bytes(N) + 2048; }
inline size_t render_min_bytes(int N, int /*n_mics*/) { return render_fixed_bytes(N) + render_row_bytes(N); }
inline size_t render_full_bytes(int N, long long rows) {
  return render_fixed_bytes(N) + size_t(std::min<long long>(rows, 4096)) * render_row_bytes(N);
}

// ---- plan of one transform length: everything the renderer derives from (N, base signal) alone -----------------
// [chirp | twiddles | chirp spectrum | X = fft(base zero-padded, 2N)].  With random rooms nearly every scene of a
// batch has its own N, and building these tables is about half of the GPU work of a small bucket: a caller that
// renders batch after batch keeps the plans (pal_render_plan) and pays for each N once.
inline size_t render_plan_bytes(int N) {
  GenericLayout<float> L(2 * N);
  return L.tables + al(sizeof(cpxf) * size_t(2 * N)) + 256;
}
inline size_t render_plan_scratch_bytes(int N) {
  GenericLayout<float> L(2 * N);
  return al(sizeof(cpxf) * size_t(L.p.M));
}
struct RenderPlan {
  BluePlan p;              // first-generation engine (power-of-two convolution length)
  BlueBuffers<float> bb;
  f2h::Buffers b2;         // second-generation engine (pal_fft2.cuh) when one of its plans holds 4N - 1 points
  cpxf* X;
  BluePlan seen() const { return b2.plan >= 0 ? b2.p : p; }       // what loaders / storers look at
};
inline RenderPlan carve_render_plan(int N, char* mem) {
  RenderPlan rp;
  const GenericLayout<float> L(2 * N);
  rp.p = L.p;
  char* b = mem;
  const int plan2 = use_fft2() ? fft2::choose_plan(2 * N) : -1;
  if (plan2 >= 0) {
    f2h::carve(2 * N, plan2, b, rp.b2);
    rp.bb.chirp = rp.b2.chirp;
    b = mem + L.tables;          // same budget as the first-generation tables (never smaller)
  } else {
    carve_plan<float>(rp.p, b, rp.bb);
  }
  rp.X = reinterpret_cast<cpxf*>(b);
  return rp;
}
// fill the plan at `mem`; `conv` is scratch for one convolution (M complex)
inline cudaError_t build_render_plan(const float* base, int n_base, int N, char* mem, cpxf* conv, cudaStream_t s, int sms) {
  using T = float;
  const RenderPlan rp = carve_render_plan(N, mem);
  if (rp.b2.plan >= 0) {
    cudaError_t e2 = f2h::fill(rp.b2, conv, s, sms);
    if (e2 != cudaSuccess) return e2;
    const BluePlan pl = rp.seen();
    // X = fft(base zero-padded, 2N)                                    (signal_processing.py:69)
    f2h::conv<false>(rp.b2, LoadSignal<T>{pl, rp.b2.chirp, base, n_base}, StoreSpectrum<T>{pl, rp.b2.chirp, rp.X}, 1, conv, s, sms);
    return cudaGetLastError();
  }
  const BluePlan& p = rp.p;
  plan_kernel_attributes<T>(p);
  cudaError_t e = fill_plan<T>(p, rp.bb, s, sms);
  if (e != cudaSuccess) return e;
  const size_t cs = col_smem<T>(p), rs = row_smem<T>(p);
  const int tiles = p.M2 / std::min(p.M2, ColTile<T>::TC);
  const BlueTables<T> tb = rp.bb.tb();
  // X = fft(base zero-padded, 2N)                                      (signal_processing.py:69)
  k_colpass_fwd<T, LoadSignal<T>><<<std::min(tiles, 16 * sms), kGT, cs, s>>>(
      p, tb, LoadSignal<T>{p, rp.bb.chirp, base, n_base}, 1, nullptr, conv);
  launch_rowpass<T, true, false>(p, tb, 1, conv, s, 16LL * sms);
  k_colpass_inv<T, StoreSpectrum<T>><<<std::min(tiles, 16 * sms), kGT, cs, s>>>(p, tb, StoreSpectrum<T>{p, rp.bb.chirp, rp.X}, 1,
                                                                                nullptr, conv);
  count_launch(3);
  return cudaGetLastError();
}

// workspace of the row loop alone (plan held elsewhere): per row in flight G[N+1], a convolution buffer, a live flag
inline size_t render_rows_min_bytes(int N) { return render_row_bytes(N) + 1024; }

// Render `n_rows` rows (bucket-local order, see RenderRows) of transform length 2N into `out`
// (row r of the batch at out + r * n_keep) with a ready plan.  No normalisation here.
inline cudaError_t render_rows_planned(const RenderPlan& rp, int N, RenderRows rr, long long n_rows, double fs, int n_keep,
                                       float* out, char* ws, size_t ws_bytes, cudaStream_t s, int sms) {
  using T = float;
  if (ws_bytes < render_rows_min_bytes(N)) return cudaErrorMemoryAllocation;
  const BluePlan& p = rp.p;
  char* b = ws;
  const size_t g_one = al(sizeof(cpxf) * size_t(N + 1)), conv_one = al(sizeof(cpxf) * size_t(p.M));
  // rows in flight (kept even: two rows share a convolution buffer, which this sizing over-provisions)
  long long cap = std::max<long long>(1, std::min<long long>(n_rows, (long long)((ws_bytes - 512) / (g_one + conv_one + sizeof(int)))));
  cap = std::max<long long>(1, std::min<long long>(cap, conv_chunk_bytes() / (long long)conv_one));   // chunk stays in L2
  if (cap > 1) cap &= ~1LL;
  cpxf* G = reinterpret_cast<cpxf*>(b);          // rows addressed densely: G[t * (N+1)]
  b += size_t(cap) * g_one;
  cpxf* conv = reinterpret_cast<cpxf*>(b);       // conv[t * M]
  b += size_t(cap) * conv_one;
  int* live = reinterpret_cast<int*>(b);         // live[row of the chunk]
  const size_t cs = col_smem<T>(p), rs = row_smem<T>(p);
  const int tiles = p.M2 / std::min(p.M2, ColTile<T>::TC);
  const BlueTables<T> tb = rp.bb.tb();
  if (rp.b2.plan < 0) {
    plan_kernel_attributes<T>(p);
    grow_smem(reinterpret_cast<const void*>(k_colpass_fwd<T, LoadHermitian2<T>>), cs);
    grow_smem(reinterpret_cast<const void*>(k_colpass_inv<T, StoreRender2<T>>), cs);
  }
  const int xtiles = (N + 1 + kGT * kXferJ - 1) / (kGT * kXferJ);
  const int kcap = rr.k_stride;
  const size_t ts = 4 * size_t((kcap + 3) & ~3) + 16 * size_t(kcap) + 16;
  if (ts > 48 * 1024) grow_smem(reinterpret_cast<const void*>(k_transfer), ts);
  const int fade = int(0.01 * N);
  for (long long r0 = 0; r0 < n_rows; r0 += cap) {
    const long long nt = std::min<long long>(cap, n_rows - r0);
    // G = X * H                                                        (main.py:104-118)
    k_transfer<<<(unsigned)std::min<long long>(nt * xtiles, 32LL * sms), kGT, ts, s>>>(rp.X, N, rr, r0, nt, fs, G, live);
    // two rows per inverse transform (LoadHermitian2)
    const long long ntr = (nt + 1) / 2;
    if (rp.b2.plan >= 0) {
      const BluePlan pl = rp.seen();
      f2h::conv<true>(rp.b2, LoadHermitian2<T>{pl, rp.b2.chirp, G, N, nt},
                      StoreRender2<T>{pl, rp.b2.chirp, out, N, n_keep, fade, rr, r0, nt, live}, ntr, conv, s, sms);
      count_launch(1);
      continue;
    }
    k_colpass_fwd<T, LoadHermitian2<T>><<<(unsigned)std::min<long long>(ntr * tiles, 16LL * sms), kGT, cs, s>>>(
        p, tb, LoadHermitian2<T>{p, rp.bb.chirp, G, N, nt}, ntr, nullptr, conv);
    launch_rowpass<T, true, true>(p, tb, ntr, conv, s, 16LL * sms);
    k_colpass_inv<T, StoreRender2<T>><<<(unsigned)std::min<long long>(ntr * tiles, 16LL * sms), kGT, cs, s>>>(
        p, tb, StoreRender2<T>{p, rp.bb.chirp, out, N, n_keep, fade, rr, r0, nt, live}, ntr, nullptr, conv);
    count_launch(4);
  }
  return cudaGetLastError();
}

// one-shot form: the plan is built at the head of the workspace, the row loop uses the rest
inline cudaError_t render_rows(const float* base, int n_base, int N, RenderRows rr, long long n_rows, double fs, int n_keep,
                               float* out, char* ws, size_t ws_bytes, cudaStream_t s, int sms) {
  if (ws_bytes < render_min_bytes(N, 1)) return cudaErrorMemoryAllocation;
  const size_t plan = (render_plan_bytes(N) + 255) / 256 * 256;
  char* rest = ws + plan;
  // the head of the row loop's workspace (at least one G row + one convolution buffer) doubles as the scratch of
  // the plan build: the build is stream-ordered before the first kernel that writes there
  cudaError_t e = build_render_plan(base, n_base, N, ws, reinterpret_cast<cpxf*>(rest), s, sms);
  if (e != cudaSuccess) return e;
  return render_rows_planned(carve_render_plan(N, ws), N, rr, n_rows, fs, n_keep, out, rest, ws_bytes - plan, s, sms);
}

// ---------------------------------------------------------------- grouped form: many buckets (transform lengths) per launch
__global__ void __launch_bounds__(kGT) k_transfer_group(RenderGroupArgs a) {
  extern __shared__ __align__(16) char smem[];
  transfer_group_body<kGT, kXferJ>(a, smem);
}
template <class P> __global__ void __launch_bounds__(f2h::kNT2, PAL_FFT2_MINBLOCKS) k_render_group_colfwd(RenderGroupArgs a) {
  extern __shared__ __align__(128) char smem[];
  render_group_colfwd_body<P, f2h::kNT2>(a, smem);
}
template <class P> __global__ void __launch_bounds__(f2h::kNT2, PAL_FFT2_MINBLOCKS) k_render_group_row(RenderGroupArgs a) {
  extern __shared__ __align__(128) char smem[];
  render_group_row_body<P, f2h::kNT2>(a, smem);
}
template <class P> __global__ void __launch_bounds__(f2h::kNT2, PAL_FFT2_MINBLOCKS) k_render_group_colinv(RenderGroupArgs a) {
  extern __shared__ __align__(128) char smem[];
  render_group_colinv_body<P, f2h::kNT2>(a, smem);
}

// pinned staging for the bucket tables (host -> device, asynchronous): a small ring per process, a slot is reused only
// after the copy that read it has completed
struct StagingRing {
  static constexpr int kSlots = 64;
  void* host[kSlots] = {};
  size_t cap[kSlots] = {};
  cudaEvent_t ev[kSlots] = {};
  int next = 0;
  void* acquire(size_t bytes, int& slot) {
    slot = next;
    next = (next + 1) % kSlots;
    if (ev[slot]) cudaEventSynchronize(ev[slot]);
    else cudaEventCreateWithFlags(&ev[slot], cudaEventDisableTiming);
    if (cap[slot] < bytes) {
      if (host[slot]) cudaFreeHost(host[slot]);
      cap[slot] = std::max<size_t>(bytes, 1 << 18);
      if (cudaHostAlloc(&host[slot], cap[slot], cudaHostAllocDefault) != cudaSuccess) { host[slot] = nullptr; cap[slot] = 0; }
    }
    return host[slot];
  }
};
inline StagingRing& staging_ring() {
  static thread_local StagingRing r;
  return r;
}

struct GroupedBucketIn {
  const char* plan;        // device memory of the bucket's render plan (pal_render_plan)
  int N;
  long long first, count;  // the bucket's scenes: scene_index[first .. first + count)
};
inline size_t grouped_bucket_bytes(int N, long long rows, int plan2) {
  const fft2::PlanDims d = fft2::plan_dims(plan2);
  return al(sizeof(cpxf) * size_t(rows) * (N + 1)) + al(sizeof(cpxf) * size_t((rows + 1) / 2) * d.M1 * d.M2) + al(sizeof(int) * size_t(rows));
}
// Render every bucket of `in` (all planned, all on the second-generation engine) into `out`; buckets that share a
// convolution plan are issued together, as many per group as the workspace holds.  cudaErrorNotSupported: a bucket has no
// second-generation plan or does not fit the workspace on its own (the caller renders that batch bucket by bucket).
inline cudaError_t render_grouped(const GroupedBucketIn* in, int n_in, RenderRows rr_all, const long long* scene_index_dev, double fs,
                                  int n_keep, float* out, char* ws, size_t ws_bytes, cudaStream_t s, int sms) {
  if (!use_fft2()) return cudaErrorNotSupported;
  const int kcap = rr_all.k_stride;
  const size_t ts = 4 * size_t((kcap + 3) & ~3) + 16 * size_t(kcap) + 16;
  if (ts > 48 * 1024) grow_smem(reinterpret_cast<const void*>(k_transfer_group), ts);
  std::vector<int> plan_of(n_in);
  for (int i = 0; i < n_in; ++i) {
    plan_of[i] = fft2::choose_plan(2 * in[i].N);
    if (plan_of[i] < 0) return cudaErrorNotSupported;
    // short transforms run as one CTA per convolution (their plans keep the chirp spectrum in that kernel's order)
    if (f2h::plan_fits_smem(plan_of[i])) return cudaErrorNotSupported;
    if (grouped_bucket_bytes(in[i].N, in[i].count * rr_all.n_mics, plan_of[i]) + (1 << 16) > ws_bytes) return cudaErrorNotSupported;
  }
  std::vector<RenderBucket> tab;
  int i0 = 0;
  while (i0 < n_in) {
    // group = maximal run i0 .. i1-1 of one plan that fits the workspace
    const int plan2 = plan_of[i0];
    const fft2::PlanDims d = fft2::plan_dims(plan2);
    size_t used = 0;
    int i1 = i0;
    while (i1 < n_in && plan_of[i1] == plan2 && i1 - i0 < 8192) {
      const size_t need = grouped_bucket_bytes(in[i1].N, in[i1].count * rr_all.n_mics, plan2);
      const size_t table = al(sizeof(RenderBucket) * size_t(i1 - i0 + 2));
      if (used + need + table + 4096 > ws_bytes) break;
      used += need;
      ++i1;
    }
    const int nb = i1 - i0;
    char* base = ws;
    RenderBucket* tab_dev = reinterpret_cast<RenderBucket*>(base);
    base += al(sizeof(RenderBucket) * size_t(nb + 1));
    tab.assign(nb + 1, RenderBucket{});
    long long xfer = 0, col = 0, row = 0;
    int tc = 16, tr = 16;
    fft2::with_plan(plan2, [&](auto pl) { tc = decltype(pl)::TC; tr = decltype(pl)::TR; });
    for (int k = 0; k < nb; ++k) {
      const GroupedBucketIn& bi = in[i0 + k];
      const RenderPlan rp = carve_render_plan(bi.N, const_cast<char*>(bi.plan));
      RenderBucket& b = tab[k];
      b.tb = rp.b2.tb();
      b.X = rp.X;
      b.scene_index = scene_index_dev + bi.first;
      b.n_rows = bi.count * rr_all.n_mics;
      b.N = bi.N;
      b.fade = int(0.01 * bi.N);
      b.G = reinterpret_cast<cpxf*>(base);    base += al(sizeof(cpxf) * size_t(b.n_rows) * (bi.N + 1));
      b.conv = reinterpret_cast<cpxf*>(base); base += al(sizeof(cpxf) * size_t((b.n_rows + 1) / 2) * d.M1 * d.M2);
      b.live = reinterpret_cast<int*>(base);  base += al(sizeof(int) * size_t(b.n_rows));
      b.xfer0 = xfer; b.col0 = col; b.row0 = row;
      const long long ntr = (b.n_rows + 1) / 2;
      xfer += b.n_rows * ((bi.N + 1 + kGT * kXferJ - 1) / (kGT * kXferJ));
      col += ntr * (d.M2 / tc);
      row += ntr * (d.M1 / tr);
    }
    tab[nb].xfer0 = xfer; tab[nb].col0 = col; tab[nb].row0 = row;      // sentinel: unit totals
    int slot = 0;
    void* host = staging_ring().acquire(sizeof(RenderBucket) * size_t(nb + 1), slot);
    if (!host) return cudaErrorMemoryAllocation;
    std::memcpy(host, tab.data(), sizeof(RenderBucket) * size_t(nb + 1));
    cudaError_t e = cudaMemcpyAsync(tab_dev, host, sizeof(RenderBucket) * size_t(nb + 1), cudaMemcpyHostToDevice, s);
    if (e != cudaSuccess) return e;
    cudaEventRecord(staging_ring().ev[slot], s);
    RenderGroupArgs a{tab_dev, nb, rr_all, fs, out, n_keep};
    a.rr.scene_index = nullptr;
    k_transfer_group<<<(unsigned)std::min<long long>(xfer, 32LL * sms), kGT, ts, s>>>(a);
    fft2::with_plan(plan2, [&](auto pl) {
      using P = decltype(pl);
      k_render_group_colfwd<P><<<f2h::grid_for<k_render_group_colfwd<P>>(P::col_smem, col, sms), f2h::kNT2, P::col_smem, s>>>(a);
      k_render_group_row<P><<<f2h::grid_for<k_render_group_row<P>>(P::row_smem, row, sms), f2h::kNT2, P::row_smem, s>>>(a);
      k_render_group_colinv<P><<<f2h::grid_for<k_render_group_colinv<P>>(P::col_smem, col, sms), f2h::kNT2, P::col_smem, s>>>(a);
    });
    count_launch(4);
    e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    i0 = i1;
  }
  return cudaSuccess;
}

inline cudaError_t normalise_rows(float* out, long long n_rows, int n_keep, cudaStream_t s, int sms) {
  k_normalise_compress<<<(unsigned)std::min<long long>(n_rows, 8LL * sms), kGT, 64, s>>>(out, n_rows, n_keep, 0.8f, 1e-8f, true);
  count_launch(1);
  return cudaGetLastError();
}

}  // namespace palhost
