// pal_render_host.cuh -- host-side orchestration of stage 1 (image sources, multipath renderer).
// Included by pal_capi.cu only.
#pragma once
#include "pal_generic_host.cuh"
#include "pal_render.cuh"

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
    cudaFuncSetAttribute(k_colpass_fwd<T, LoadHermitian2<T>>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cs);
    cudaFuncSetAttribute(k_colpass_inv<T, StoreRender2<T>>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cs);
  }
  const int xtiles = (N + 1 + kGT * kXferJ - 1) / (kGT * kXferJ);
  const int kcap = rr.k_stride;
  const size_t ts = 4 * size_t((kcap + 3) & ~3) + 16 * size_t(kcap) + 16;
  if (ts > 48 * 1024) cudaFuncSetAttribute(k_transfer, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ts);
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

inline cudaError_t normalise_rows(float* out, long long n_rows, int n_keep, cudaStream_t s, int sms) {
  k_normalise_compress<<<(unsigned)std::min<long long>(n_rows, 8LL * sms), kGT, 64, s>>>(out, n_rows, n_keep, 0.8f, 1e-8f, true);
  count_launch(1);
  return cudaGetLastError();
}

}  // namespace palhost
