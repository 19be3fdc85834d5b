// pal_render.cuh -- stage 1 of the hot path: image sources and the multipath renderer.
//
//   image_sources_body   utils.py:67-106 (+ :29-65)  breadth-first image sources, one block per scene
//   path_table_body      main.py:94-116              per (mic, path) delay and gain
//   transfer_body        main.py:104-118 + signal_processing.py:66-73: because every delayed copy shares one
//                        FFT and one fade window, sum_k a_k * fractional_delay(x, tau_k) equals
//                        w * irfft(rfft(x, 2N) * H), H[m] = sum_k a_k exp(-j 2 pi m tau_k fs / (2N));
//                        this kernel accumulates H for a tile of bins (no atomics) and multiplies by X
//   LoadHermitian / StoreRender   the length-2N inverse DFT runs through the Bluestein kernels
//   normalise_compress_body       main.py:119-122, signal_processing.py:82-94
#pragma once
#include "pal_bluestein.cuh"
#include "pal_fft2.cuh"

namespace pal {

// ------------------------------------------------------------------ image sources
#if PAL_GPU
PAL_DEV double dmul(double a, double b) { return __dmul_rn(a, b); }   // no FMA contraction: the reference
PAL_DEV double dadd(double a, double b) { return __dadd_rn(a, b); }   // evaluates with python scalars
PAL_DEV double ddiv(double a, double b) { return __ddiv_rn(a, b); }
PAL_DEV long long round_key(double x, double scale) { return __double2ll_rn(__dmul_rn(x, scale)); }
#else
inline double dmul(double a, double b) { volatile double r = a * b; return r; }
inline double dadd(double a, double b) { volatile double r = a + b; return r; }
inline double ddiv(double a, double b) { volatile double r = a / b; return r; }
inline long long round_key(double x, double scale) { return (long long)std::nearbyint(dmul(x, scale)); }
#endif

// np.linalg.norm of a 3-vector (utils.py:44-48) is sqrt(x.dot(x)); the BLAS dot of three elements evaluates
// fma(z, z, fma(y, y, x*x)) (checked against numpy on 20000 random vectors, tests/test_oracle.py), so the distance
// is formed with exactly these roundings -- a prune decision at a 1-ulp tie then falls the same way.
#if PAL_GPU
PAL_DEV double norm3(double x, double y, double z) { return __dsqrt_rn(__fma_rn(z, z, __fma_rn(y, y, __dmul_rn(x, x)))); }
#else
inline double norm3(double x, double y, double z) { return std::sqrt(std::fma(z, z, std::fma(y, y, dmul(x, x)))); }
#endif

struct ImgParams {
  int n_planes, n_mics, max_order, k_max;
  double frequency, threshold, round_scale;
};

// utils.py:50-65 for one distance
PAL_DEV double attenuation(double d, double absorption, double freq_factor, double frequency) {
  d = d < 0.1 ? 0.1 : d;
  return (1.0 / d) * exp(-freq_factor * frequency * d) * exp(-absorption * d);
}

// numpy's add.reduce over a contiguous float64 array (what np.mean of a list does): fewer than 8 values are summed
// left to right; up to 128 values go through eight interleaved accumulators combined as
// ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7)) with the tail added one by one; longer arrays are split in halves (the first
// a multiple of 8) and the two partial sums added.  f(q) yields value q.
template <class F> PAL_DEV double numpy_pairwise_sum(F& f, int lo, int n) {
  if (n < 8) {
    double r = 0.0;
    for (int i = 0; i < n; ++i) r = dadd(r, f(lo + i));
    return r;
  }
  if (n <= 128) {
    double r[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) r[j] = f(lo + j);
    int i = 8;
    for (; i < n - (n % 8); i += 8) {
#pragma unroll
      for (int j = 0; j < 8; ++j) r[j] = dadd(r[j], f(lo + i + j));
    }
    double res = dadd(dadd(dadd(r[0], r[1]), dadd(r[2], r[3])), dadd(dadd(r[4], r[5]), dadd(r[6], r[7])));
    for (; i < n; ++i) res = dadd(res, f(lo + i));
    return res;
  }
  int n2 = n / 2;
  n2 -= n2 % 8;
  const double a = numpy_pairwise_sum(f, lo, n2);
  return dadd(a, numpy_pairwise_sum(f, lo + n2, n - n2));
}

// One block per scene (grid-stride).  Scratch per BLOCK: cand_pos[cmax][3] f64, cand_key[cmax][3] i64,
// cand_ok[cmax] i32 with cmax = k_max * n_planes; keys[k_max+1][3] i64.
// Output per scene: pos[k_max][3], mat[k_max] (plane material id), count (or -1 on overflow).
template <int NT>
PAL_DEV void image_sources_body(ImgParams ip, const double* sources, long long n_scenes, const double* planes_all,
                                long long plane_stride /* 0: shared planes */, const int* plane_mat, const double* mat_abs, const double* mat_freq,
                                const double* mics, long long mic_stride /* 0: shared array */, double* out_pos,
                                int* out_mat, int* out_count, char* scratch_all, size_t scratch_per_block,
                                char* smem_raw) {
  int* sh = reinterpret_cast<int*>(smem_raw);   // [0] running offset, [1..NT/32] warp sums
  const int tid = simt::tid();
  const int cmax = ip.k_max * ip.n_planes;
  char* sc = scratch_all + size_t(simt::bid()) * scratch_per_block;
  double* cand_pos = reinterpret_cast<double*>(sc);
  long long* cand_key = reinterpret_cast<long long*>(cand_pos + size_t(cmax) * 3);
  long long* keys = cand_key + size_t(cmax) * 3;
  int* cand_ok = reinterpret_cast<int*>(keys + size_t(ip.k_max + 1) * 3);
  for (long long s = simt::bid(); s < n_scenes; s += simt::nblocks()) {
    const double* src = sources + s * 3;
    const double* mic = mics + s * mic_stride;
    const double* planes = planes_all + s * plane_stride;
    double* pos = out_pos + s * ip.k_max * 3;
    int* mat = out_mat + s * ip.k_max;
    if (tid < 3) keys[tid] = round_key(src[tid], ip.round_scale);      // the source itself is "seen"
    simt::sync_block();
    int n_img = 0, lvl0 = -1, lvl1 = 0;    // frontier = images [lvl0, lvl1); lvl0 == -1: the source
    bool overflow = false;
    for (int order = 1; order <= ip.max_order; ++order) {
      const int nf = (lvl0 < 0) ? 1 : (lvl1 - lvl0);
      const int nc = nf * ip.n_planes;
      if (nc == 0) break;
      // 1. candidates: reflect, key, prune test (utils.py:89-99)
      for (int c = tid; c < nc; c += NT) {
        const int par = c / ip.n_planes, pl = c % ip.n_planes;
        const double* p = (lvl0 < 0) ? src : pos + size_t(lvl0 + par) * 3;
        const double a = planes[pl * 4], b = planes[pl * 4 + 1], cc = planes[pl * 4 + 2], d = planes[pl * 4 + 3];
        const double den = dadd(dadd(dmul(a, a), dmul(b, b)), dmul(cc, cc));
        const double num = dadd(dadd(dadd(dmul(a, p[0]), dmul(b, p[1])), dmul(cc, p[2])), d);
        const double f = ddiv(dmul(2.0, num), den);
        const double x = dadd(p[0], -dmul(a, f)), y = dadd(p[1], -dmul(b, f)), z = dadd(p[2], -dmul(cc, f));
        cand_pos[c * 3] = x; cand_pos[c * 3 + 1] = y; cand_pos[c * 3 + 2] = z;
        cand_key[c * 3] = round_key(x, ip.round_scale);
        cand_key[c * 3 + 1] = round_key(y, ip.round_scale);
        cand_key[c * 3 + 2] = round_key(z, ip.round_scale);
        const int m = plane_mat[pl];
        double mn = 1e300;
        auto att_of = [&](int q) {
          const double att = attenuation(norm3(dadd(x, -mic[q * 3]), dadd(y, -mic[q * 3 + 1]), dadd(z, -mic[q * 3 + 2])),
                                         mat_abs[m], mat_freq[m], ip.frequency);
          mn = att < mn ? att : mn;
          return att;
        };
        const double sum = numpy_pairwise_sum(att_of, 0, ip.n_mics);        // np.mean(attenuations), utils.py:99
        cand_ok[c] = (ddiv(sum, double(ip.n_mics)) > ip.threshold && mn > ip.threshold / 2) ? 1 : 0;
      }
      simt::sync_block();
      // 2. a passing candidate is new iff its key is neither among the seen keys (source + accepted
      //    images of earlier levels) nor carried by an EARLIER PASSING candidate of this level
      //    (failed candidates are not added to `seen`, utils.py:99-100)
      for (int c = tid; c < nc; c += NT) {
        if (!cand_ok[c]) continue;
        const long long k0 = cand_key[c * 3], k1 = cand_key[c * 3 + 1], k2 = cand_key[c * 3 + 2];
        bool dup = false;
        for (int j = 0; j <= n_img && !dup; ++j) dup = keys[j * 3] == k0 && keys[j * 3 + 1] == k1 && keys[j * 3 + 2] == k2;
        for (int j = 0; j < c && !dup; ++j)
          dup = cand_ok[j] && cand_key[j * 3] == k0 && cand_key[j * 3 + 1] == k1 && cand_key[j * 3 + 2] == k2;
        if (dup) cand_ok[c] = 2;    // 2 = passing duplicate (keeps masking later twins, but is not accepted)
      }
      simt::sync_block();
      // 3. append the accepted candidates in candidate order (block-wide exclusive scan)
      if (tid == 0) sh[0] = n_img;
      simt::sync_block();
      for (int c0 = 0; c0 < nc; c0 += NT) {
        const int c = c0 + tid;
        const int acc = (c < nc && cand_ok[c] == 1) ? 1 : 0;
        int incl = acc;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const int v = simt::shfl(incl, simt::lane() - o);
          if (simt::lane() >= o) incl += v;
        }
        if (simt::lane() == 31) sh[1 + simt::warp()] = incl;
        simt::sync_block();
        int woff = 0;
        for (int w = 0; w < simt::warp(); ++w) woff += sh[1 + w];
        int tot = 0;
        for (int w = 0; w < NT / 32; ++w) tot += sh[1 + w];
        const int base = sh[0];
        const int slot = base + woff + incl - acc;
        if (acc) {
          if (slot < ip.k_max) {
            pos[slot * 3] = cand_pos[c * 3]; pos[slot * 3 + 1] = cand_pos[c * 3 + 1]; pos[slot * 3 + 2] = cand_pos[c * 3 + 2];
            mat[slot] = plane_mat[c % ip.n_planes];
            keys[(slot + 1) * 3] = cand_key[c * 3]; keys[(slot + 1) * 3 + 1] = cand_key[c * 3 + 1];
            keys[(slot + 1) * 3 + 2] = cand_key[c * 3 + 2];
          }
        }
        simt::sync_block();
        if (tid == 0) sh[0] = base + tot;
        simt::sync_block();
      }
      const int new_total = sh[0];
      simt::sync_block();
      if (new_total > ip.k_max) { overflow = true; break; }
      lvl0 = n_img;
      lvl1 = new_total;
      n_img = new_total;
      if (lvl1 == lvl0) break;
    }
    if (tid == 0) out_count[s] = overflow ? -1 : n_img;
    simt::sync_block();
  }
}

// ------------------------------------------------------------------ path table
// thread per (mic, path) of one scene; path 0 = direct (material 'air', main.py:108), path k = image
// k-1.  tau [M][K1] f64 seconds; gain [M][K1] f64 RAW attenuation evaluated like the reference
// (utils.py:50-65; it may be 1e-38, denormal or exactly 0 -- the renderer divides by the per-mic
// maximum in float64 before anything is rounded to fp32, and a zero maximum yields a zero row).
PAL_DEV void path_table_body(const double* src, const double* img_pos, const int* img_mat, int n_img,
                             const double* mics, int n_mics, const double* mat_abs, const double* mat_freq,
                             int air_mat, double frequency, double c_sound, double* tau, double* gain) {
  const int k1 = n_img + 1;
  const long long i = (long long)simt::bid() * simt::nthreads() + simt::tid();
  if (i >= (long long)n_mics * k1) return;
  const int m = int(i / k1), k = int(i % k1);
  const double* p = (k == 0) ? src : img_pos + size_t(k - 1) * 3;
  const int mat = (k == 0) ? air_mat : img_mat[k - 1];
  const double dx = p[0] - mics[m * 3], dy = p[1] - mics[m * 3 + 1], dz = p[2] - mics[m * 3 + 2];
  const double d = norm3(dx, dy, dz);
  tau[i] = d / c_sound;
  gain[i] = attenuation(d, mat_abs[mat], mat_freq[mat], frequency);
}

// Batched form: one block per scene (grid-stride).  tau / gain [S][M][k_stride] with k_stride >= count+1,
// path_count[s] = img_count[s] + 1 (0 when the image list overflowed), max_tau[s] = max over mics and
// paths of the delay (main.py:94-101), reduced inside the block -- no atomics.
template <int NT>
PAL_DEV void path_table_batched_body(const double* sources, const double* img_pos, const int* img_mat,
                                     const int* img_count, long long n_scenes, int k_max, const double* mics,
                                     int n_mics, long long mic_stride, const double* mat_abs, const double* mat_freq,
                                     int air_mat, double frequency, double c_sound, int k_stride, double* tau,
                                     double* gain, int* path_count, double* max_tau, char* smem_raw) {
  double* sh = reinterpret_cast<double*>(smem_raw);   // [NT/32]
  for (long long s = simt::bid(); s < n_scenes; s += simt::nblocks()) {
    const int cnt = img_count[s];
    const int k1 = (cnt < 0) ? 0 : cnt + 1;
    const double* src = sources + s * 3;
    const double* mic = mics + s * mic_stride;
    double mx = 0.0;
    for (int i = simt::tid(); i < n_mics * k1; i += NT) {
      const int m = i / k1, k = i % k1;
      const double* p = (k == 0) ? src : img_pos + (size_t(s) * k_max + (k - 1)) * 3;
      const int mat = (k == 0) ? air_mat : img_mat[size_t(s) * k_max + (k - 1)];
      const double dx = p[0] - mic[m * 3], dy = p[1] - mic[m * 3 + 1], dz = p[2] - mic[m * 3 + 2];
      const double d = norm3(dx, dy, dz);
      const double t = d / c_sound;
      const size_t o = (size_t(s) * n_mics + m) * k_stride + k;
      tau[o] = t;
      gain[o] = attenuation(d, mat_abs[mat], mat_freq[mat], frequency);
      mx = t > mx ? t : mx;
    }
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) {
      const double o = simt::shfl_xor(mx, m);
      mx = o > mx ? o : mx;
    }
    if (simt::lane() == 0) sh[simt::warp()] = mx;
    simt::sync_block();
    if (simt::tid() == 0) {
      double r = 0.0;
      for (int w = 0; w < NT / 32; ++w) r = sh[w] > r ? sh[w] : r;
      max_tau[s] = r;
      path_count[s] = k1;
    }
    simt::sync_block();
  }
}

// Which (scene, mic) a renderer row is, and where its path list lives.  A "local" row is a row of
// the bucket being rendered (scenes that share the transform length N); scene_index maps bucket
// scenes to scenes of the batch.
struct RenderRows {
  const double* tau;             // [scenes][n_mics][k_stride]
  const double* gain;
  const int* path_count;         // [scenes] paths per row (direct + images); nullptr: k_stride for all
  const long long* scene_index;  // [bucket scenes] -> scene; nullptr: identity
  int k_stride, n_mics;
  PAL_DEV long long global_row(long long local_row) const {
    const long long sl = local_row / n_mics;
    return (scene_index ? scene_index[sl] : sl) * n_mics + local_row % n_mics;
  }
  PAL_DEV int paths(long long global_row_) const { return path_count ? path_count[global_row_ / n_mics] : k_stride; }
};

// ------------------------------------------------------------------ transfer function x spectrum
// G[mic][m] = X[m] * sum_k g_k exp(-j 2 pi m tau_k fs / (2N)),  m = 0..N   (bin N: real part only,
// sum_k g_k cos(pi fs tau_k), because the reference's fftfreq labels it -fs/2 and keeps .real)
// One block per (mic, tile of NT*J bins); each thread owns bins m0 + t + NT*j and advances its
// phasor by a per-path rotation of NT bins; phases are reduced mod 1 in float64 before any
// single-precision trigonometry (SURVEY.md hard part 6).
// one work unit: row `lrow` of the chunk (G / live are the chunk's buffers), bins m0 .. m0 + NT*J - 1
template <int NT, int J>
PAL_DEV void transfer_unit(const cpxf* X, int N, const RenderRows& rr, long long row0, long long lrow, int m0, double fs, cpxf* G,
                           int* live /* [n_rows]: 0 = the row has no audible path, its output is exactly zero */, char* smem_raw) {
  const int kcap = rr.k_stride;
  float* s_gain = reinterpret_cast<float*>(smem_raw);         // [k1]
  cpxf* s_rot = reinterpret_cast<cpxf*>(s_gain + ((kcap + 3) & ~3));   // [k1] rotation by NT bins
  double* s_delta = reinterpret_cast<double*>(s_rot + kcap);   // [k1] turns per bin
  const int tid = simt::tid();
  const int nbins = N + 1;
  {
    const long long grow = rr.global_row(row0 + lrow);
    const int k1 = rr.paths(grow);
    const double* tk = rr.tau + size_t(grow) * rr.k_stride;
    const double* gk = rr.gain + size_t(grow) * rr.k_stride;
    double gmx = 0.0;
    for (int k = 0; k < k1; ++k) gmx = gk[k] > gmx ? gk[k] : gmx;        // small k1, L1-resident
    if (live && m0 == 0 && tid == 0) live[lrow] = (k1 > 0 && gmx > 0.0) ? 1 : 0;
    for (int k = tid; k < k1; k += NT) {
      const double delta = tk[k] * fs / (2.0 * N);
      s_delta[k] = delta;
      s_gain[k] = (gmx > 0.0) ? float(gk[k] / gmx) : 0.f;
      double fr = delta * NT;
      fr -= floor(fr);
      float sn, cs;
      sincospi_<float>(2.0 * fr, sn, cs);
      s_rot[k] = cpxf{cs, -sn};
    }
    simt::sync_block();
    float ar[J], ai[J];
#pragma unroll
    for (int j = 0; j < J; ++j) { ar[j] = 0.f; ai[j] = 0.f; }
    const int mt = m0 + tid;
    // bins this thread really owns (the last tile of a row is partly empty: its threads stop early instead of
    // accumulating phasors nobody stores)
    const int jn = mt < nbins ? ((nbins - mt + NT - 1) / NT < J ? (nbins - mt + NT - 1) / NT : J) : 0;
    for (int k = 0; k < k1; ++k) {
      // seed phasor at bin mt: the phase is reduced mod 1 in float64 (m tau fs / 2N reaches thousands of turns,
      // SURVEY.md hard part 6); the reduced turn fraction in [0, 1) is then exact to 6e-8 in float32, which is what
      // the single-precision sincospi needs (3.7e-7 rad: far inside the 1e-5 rendering tolerance)
      double fr = s_delta[k] * double(mt);
      fr -= floor(fr);
      float sn, cs;
      sincospif_(float(2.0 * fr), sn, cs);
      float zr = cs, zi = -sn;
      const float g = s_gain[k];
      const cpxf r = s_rot[k];
#pragma unroll
      for (int j = 0; j < J; ++j) {
        if (j < jn) {
          ar[j] = fmaf(g, zr, ar[j]);
          ai[j] = fmaf(g, zi, ai[j]);
          const float nr = fmaf(zr, r.x, -(zi * r.y));
          zi = fmaf(zr, r.y, zi * r.x);
          zr = nr;
        }
      }
    }
#pragma unroll
    for (int j = 0; j < J; ++j) {
      const int m = mt + NT * j;
      if (m < nbins) {
        const cpxf x = X[m];
        cpxf o{fmaf(x.x, ar[j], -(x.y * ai[j])), fmaf(x.x, ai[j], x.y * ar[j])};
        if (m == N) {   // Nyquist bin: X[N] is real, H = sum g cos(pi fs tau)
          float h = 0.f;
          for (int k = 0; k < k1; ++k) {
            double fr = 0.5 * fs * tk[k];
            fr -= floor(fr);
            float sn, cs;
            sincospi_<float>(2.0 * fr, sn, cs);
            h = fmaf(s_gain[k], cs, h);
          }
          o = cpxf{x.x * h, 0.f};
        }
        G[size_t(lrow) * nbins + m] = o;
      }
    }
    simt::sync_block();
  }
}
template <int NT, int J>
PAL_DEV void transfer_body(const cpxf* X, int N, RenderRows rr, long long row0, long long n_rows, double fs, cpxf* G, int* live,
                           char* smem_raw) {
  const int tiles = (N + 1 + NT * J - 1) / (NT * J);
  for (long long u = simt::bid(); u < n_rows * tiles; u += simt::nblocks())
    transfer_unit<NT, J>(X, N, rr, row0, u / tiles, int(u % tiles) * NT * J, fs, G, live, smem_raw);
}

// ------------------------------------------------------------------ Bluestein loader / storer
// inverse DFT of length 2N of the Hermitian extension of G[row][0..N]
// Two rows per transform (pal_bluestein.cuh): the rendered channels are real, so rows 2t and 2t+1 of a chunk travel as
// the real and imaginary part of ONE inverse transform, a[k] = (Ga[k] + i Gb[k]) conj(chirp[k]) with Ga, Gb the
// Hermitian extensions of the half spectra.  The imaginary parts of the DC and Nyquist bins, which a real inverse
// transform ignores (and `.real` of signal_processing.py:72 discards), are dropped explicitly: they would otherwise
// leak into the partner row.
template <typename T> struct LoadHermitian2 {
  BluePlan p;              // p.n == 2N
  const cpx<T>* chirp;
  const cpxf* G;           // [rows][N+1]
  int N;
  long long n_rows;        // rows of this launch
  struct Ctx {
    const cpxf *ga, *gb;   // gb == nullptr: no second row
  };
  PAL_DEV Ctx begin(long long t) const {
    return Ctx{G + (2 * t) * (N + 1), (2 * t + 1 < n_rows) ? G + (2 * t + 1) * (N + 1) : nullptr};
  }
  PAL_DEV cpx<T> half(const cpxf* g_row, int k) const {
    const cpxf g = g_row[(k <= N) ? k : p.n - k];
    const bool self_conj = (k == 0) || (k == N);
    return cpx<T>{T(g.x), self_conj ? T(0) : ((k <= N) ? T(g.y) : T(-g.y))};
  }
  // unconditional loads (index clamped, value masked) so that the loads of several samples can be in flight together
  PAL_DEV cpx<T> operator()(const Ctx& c, int k) const {
    const bool in = k < p.n;
    const int kc = in ? k : 0;
    const T m = in ? T(1) : T(0);
    const cpx<T> a = half(c.ga, kc);
    const cpx<T> b = half(c.gb ? c.gb : c.ga, kc);
    const T mb = c.gb ? m : T(0);
    return cmulc(cpx<T>{a.x * m - b.y * mb, a.y * m + b.x * mb}, chirp[kc]);      // inverse transform: conjugate chirp
  }
};
// y[j] = Re / Im (conv[j] * conj(chirp[j])) / (2N) * fade[j], j < n_keep   (signal_processing.py:72-79)
template <typename T> struct StoreRender2 {
  BluePlan p;
  const cpx<T>* chirp;
  float* out;              // [all rows][n_keep]
  int N, n_keep, fade;
  RenderRows rr;           // chunk-local row r -> output row rr.global_row(row0 + r)
  long long row0, n_rows;
  const int* live;         // [n_rows] from the transfer kernel: a silent row must come out as exact zeros, not as the
                           // rounding residue of its partner row (normalize_signal would blow that up to full scale)
  struct Ctx {
    float *oa, *ob;        // ob == nullptr: no second row
    bool live_a, live_b;
  };
  PAL_DEV Ctx begin(long long t) const {
    Ctx c;
    c.oa = out + rr.global_row(row0 + 2 * t) * n_keep;
    c.live_a = live[2 * t] != 0;
    const bool has_b = 2 * t + 1 < n_rows;
    c.ob = has_b ? out + rr.global_row(row0 + 2 * t + 1) * n_keep : nullptr;
    c.live_b = has_b && live[2 * t + 1] != 0;
    return c;
  }
  PAL_DEV void operator()(const Ctx& c, int j, cpx<T> y) const {
    if (j >= n_keep) return;
    const cpx<T> w = chirp[j];
    T f = T(1) / T(p.n);
    if (j < fade) f *= (fade > 1) ? T(j) / T(fade - 1) : T(0);                      // np.linspace(0, 1, fade)
    if (j >= N - fade) f *= (fade > 1) ? T(N - 1 - j) / T(fade - 1) : T(1);         // np.linspace(1, 0, fade)
    c.oa[j] = c.live_a ? float(fma_(y.x, w.x, y.y * w.y) * f) : 0.f;
    if (c.ob) c.ob[j] = c.live_b ? float(fma_(y.y, w.x, -(y.x * w.y)) * f) : 0.f;
  }
};

// ------------------------------------------------------------------ many buckets per launch
// With random rooms nearly every scene has its own transform length 2N (N = int((duration + max delay) fs),
// main.py:102): a batch of 16384 scenes falls into ~1500 buckets of a dozen scenes, and one set of launches per
// bucket (16-block grids, four launches each) leaves the renderer bound by launch overhead.  A GROUP is a list of
// buckets that share the convolution plan (M1 x M2; only the chirp tables and the base spectrum differ with N): each
// of the four kernels runs ONCE per group, a block looks its work unit up in the bucket table (prefix sums of units)
// and builds the bucket's loader / storer on the fly.
struct RenderBucket {
  fft2::Tables tb;                // chirp (length 2N), stage twiddles, twf, chirp spectrum of THIS N
  const cpxf* X;                  // fft(base zero-padded, 2N)                         (signal_processing.py:69)
  const long long* scene_index;   // the bucket's scenes (indices into the batch)
  cpxf* G;                        // [n_rows][N + 1]
  cpxf* conv;                     // [ceil(n_rows / 2)][M]
  int* live;                      // [n_rows]
  long long n_rows;               // bucket scenes x microphones
  long long xfer0, col0, row0;    // first work unit of the bucket in the transfer / column / row kernels
  int N, fade;
};
template <int WHICH> PAL_DEV long long bucket_first(const RenderBucket& b) { return WHICH == 0 ? b.xfer0 : (WHICH == 1 ? b.col0 : b.row0); }
// index of the bucket that owns unit u (the table ends with a sentinel bucket holding the unit totals)
template <int WHICH> PAL_DEV int find_bucket(const RenderBucket* bk, int nb, long long u) {
  int lo = 0, hi = nb - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (bucket_first<WHICH>(bk[mid]) <= u) lo = mid; else hi = mid - 1;
  }
  return lo;
}
struct RenderGroupArgs {
  const RenderBucket* buckets;    // [n_buckets + 1] (last: sentinel with the unit totals)
  int n_buckets;
  RenderRows rr;                  // batch-wide path tables; scene_index is taken from the bucket
  double fs;
  float* out;                     // [all rows][n_keep]
  int n_keep;
};
template <int NT, int J> PAL_DEV void transfer_group_body(RenderGroupArgs a, char* smem) {
  const long long total = a.buckets[a.n_buckets].xfer0;
  for (long long u = simt::bid(); u < total; u += simt::nblocks()) {
    const RenderBucket& b = a.buckets[find_bucket<0>(a.buckets, a.n_buckets, u)];
    const int tiles = (b.N + 1 + NT * J - 1) / (NT * J);
    const long long lu = u - b.xfer0;
    RenderRows rr = a.rr;
    rr.scene_index = b.scene_index;
    transfer_unit<NT, J>(b.X, b.N, rr, 0, lu / tiles, int(lu % tiles) * NT * J, a.fs, b.G, b.live, smem);
  }
}
template <class P, int NT> PAL_DEV void render_group_colfwd_body(RenderGroupArgs a, char* smem) {
  constexpr int tiles = P::M2 / P::TC;
  f2* tile = reinterpret_cast<f2*>(smem);
  const f2* tw = fft2::stage_tw<NT>(a.buckets[0].tb.tw1, P::M1, tile + P::M1 * P::TC);      // plan-wide: any bucket's copy
  const long long total = a.buckets[a.n_buckets].col0;
  for (long long u = simt::bid(); u < total; u += simt::nblocks()) {
    const RenderBucket& b = a.buckets[find_bucket<1>(a.buckets, a.n_buckets, u)];
    const long long lu = u - b.col0;
    const BluePlan pl{2 * b.N, P::M, P::M1, P::M2, 0, 0};
    const LoadHermitian2<float> ld{pl, b.tb.chirp, b.G, b.N, b.n_rows};
    fft2::colpass_fwd_unit<P, NT>(b.tb, ld, lu / tiles, int(lu % tiles), b.conv, tile, tw);
  }
}
template <class P, int NT> PAL_DEV void render_group_row_body(RenderGroupArgs a, char* smem) {
  constexpr int tiles = P::M1 / P::TR;
  f2* tile = reinterpret_cast<f2*>(smem);
  const f2* tw = fft2::stage_tw<NT>(a.buckets[0].tb.tw2, P::M2, tile + P::M2 * P::LDR);
  const long long total = a.buckets[a.n_buckets].row0;
  for (long long u = simt::bid(); u < total; u += simt::nblocks()) {
    const RenderBucket& b = a.buckets[find_bucket<2>(a.buckets, a.n_buckets, u)];
    const long long lu = u - b.row0;
    fft2::rowpass_unit<P, NT, 1>(b.tb, lu / tiles, int(lu % tiles), b.conv, nullptr, tile, tw);
  }
}
template <class P, int NT> PAL_DEV void render_group_colinv_body(RenderGroupArgs a, char* smem) {
  constexpr int tiles = P::M2 / P::TC;
  f2* tile = reinterpret_cast<f2*>(smem);
  const f2* tw = fft2::stage_tw<NT>(a.buckets[0].tb.tw1, P::M1, tile + P::M1 * P::TC);
  const long long total = a.buckets[a.n_buckets].col0;
  for (long long u = simt::bid(); u < total; u += simt::nblocks()) {
    const RenderBucket& b = a.buckets[find_bucket<1>(a.buckets, a.n_buckets, u)];
    const long long lu = u - b.col0;
    const BluePlan pl{2 * b.N, P::M, P::M1, P::M2, 0, 0};
    RenderRows rr = a.rr;
    rr.scene_index = b.scene_index;
    const StoreRender2<float> st{pl, b.tb.chirp, a.out, b.N, a.n_keep, b.fade, rr, 0, b.n_rows, b.live};
    fft2::colpass_inv_unit<P, NT>(b.tb, st, lu / tiles, int(lu % tiles), b.conv, tile, tw);
  }
}

// ------------------------------------------------------------------ normalise + log compressor
// one block per row, in place: x / max|x| ; sign(x) * log1p(|x|/0.8 + 1e-8) / log1p(1.25 + 1e-8)
template <int NT> PAL_DEV void normalise_compress_body(float* rows, long long n_rows, int n, float threshold,
                                                       float epsilon, bool compress, char* smem_raw) {
  float* sh = reinterpret_cast<float*>(smem_raw);
  for (long long r = simt::bid(); r < n_rows; r += simt::nblocks()) {
    float* x = rows + r * n;
    float mx = 0.f;
    for (int i = simt::tid(); i < n; i += NT) mx = fmaxf(mx, fabsf(x[i]));
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) mx = fmaxf(mx, simt::shfl_xor(mx, m));
    if (simt::lane() == 0) sh[simt::warp()] = mx;
    simt::sync_block();
    mx = 0.f;
    for (int w = 0; w < NT / 32; ++w) mx = fmaxf(mx, sh[w]);
    simt::sync_block();
    if (mx > 0.f) {
      const float denom = log1pf(1.0f / threshold + epsilon);
      for (int i = simt::tid(); i < n; i += NT) {
        const float v = x[i] / mx;
        const float c = compress ? log1pf(fabsf(v) / threshold + epsilon) / denom : fabsf(v);
        x[i] = (v > 0.f) ? c : ((v < 0.f) ? -c : 0.f);
      }
    }
  }
}

}  // namespace pal
