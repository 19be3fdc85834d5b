// pal_sync.cuh -- the alignment step between the two stages: utils.synchronize_signals_improved
// (utils.py:407-457), SURVEY.md section 8f rank 2.
//
// Per scene of M channels the reference (i) picks the highest-energy channel as the alignment
// reference (:415-416), (ii) takes the FULL (non-whitened) cross-correlation of every channel with it
// (scipy.signal.correlate(sig, reference, 'full'), :426) and the arg-max of its magnitude (:427),
// (iii) refines the arg-max on a cubic spline through five correlation samples (:431-437) and
// (iv) left-pads the channels by the rounded shifts (:448-457).
//
// The device does (i), (ii) and the copy of (iv): the correlation of length len_m + len_ref - 1 is a
// circular correlation of length n = 2N-1 (N = padded row length), evaluated with the same exact
// length-n Bluestein transforms as the GCC-PHAT path (pal_bluestein.cuh) -- the loader below is
// LoadPhat without the whitening.  Everything is float64: the chain ends in round(), and the
// reference works in float64.  The five-sample spline of (iii) is 20 flops per channel and stays
// with scipy on the host, fed with the device's correlation samples.
#pragma once
#include "pal_bluestein.cuh"

namespace pal {

// forward DFT of one float64 channel row: a[j] = x[j] * chirp[j]; rows have `ld` doubles of which
// lens[row] (or `len`) are signal
template <typename T> struct LoadSignalF64 {
  BluePlan p;
  const cpx<T>* chirp;
  const double* sig;
  long long ld;
  int len;
  const int* lens;       // optional [rows]
  long long row0;
  PAL_DEV cpx<T> operator()(long long t, int j) const {
    const long long row = row0 + t;
    const int l = lens ? lens[row] : len;
    if (j >= l) return cpx<T>{T(0), T(0)};
    const T x = T(sig[row * ld + j]);
    const cpx<T> w = chirp[j];
    return cpx<T>{x * w.x, x * w.y};
  }
};

// inverse DFT of the plain cross spectrum of channel m with the scene's reference channel:
//   a[k] = S_m[k] conj(S_ref[k]) conj(chirp[k]) / n            (scipy.signal.correlate, utils.py:426)
// item = scene * Mics + m over the scenes currently resident; ref[scene0 + scene] is the reference row
template <typename T> struct LoadCross {
  BluePlan p;
  const cpx<T>* chirp;
  const cpx<T>* spec;
  const int* ref;        // [scenes] reference channel of every scene (global scene index)
  long long scene0;      // global index of resident scene 0
  int Mics;
  long long t_off;
  PAL_DEV cpx<T> operator()(long long t, int k) const {
    if (k >= p.n) return cpx<T>{T(0), T(0)};
    const long long it = t + t_off;
    const long long f = it / Mics;
    const long long ri = it, rj = f * Mics + ref[scene0 + f];
    const cpx<T> a = spec[ri * p.n + k], b = spec[rj * p.n + k];
    const T sc = T(1) / T(p.n);
    const cpx<T> x{fma_(a.x, b.x, a.y * b.y) * sc, fma_(a.y, b.x, -(a.x * b.y)) * sc};
    return cmulc(x, chirp[k]);
  }
};

// (i) energies and reference channel: one block per scene, one warp per channel at a time.
// np.argmax returns the FIRST maximum (utils.py:416).
template <int NT>
PAL_DEV void sync_energy_body(const double* sig, long long n_scenes, int Mics, long long ld, int len, const int* lens,
                              double* energy, int* ref_idx, char* smem) {
  double* se = reinterpret_cast<double*>(smem);   // [Mics]
  for (long long s = simt::bid(); s < n_scenes; s += simt::nblocks()) {
    for (int m = simt::warp(); m < Mics; m += NT / 32) {
      const long long row = s * Mics + m;
      const int l = lens ? lens[row] : len;
      const double* x = sig + row * ld;
      double acc = 0.0;
      for (int j = simt::lane(); j < l; j += 32) acc = fma_(x[j], x[j], acc);
      acc = warp_sum(acc);
      if (simt::lane() == 0) se[m] = acc;
    }
    simt::sync_block();
    if (simt::tid() == 0) {
      int best = 0;
      for (int m = 0; m < Mics; ++m) {
        if (energy) energy[s * Mics + m] = se[m];
        if (se[m] > se[best]) best = m;
      }
      ref_idx[s] = best;
    }
    simt::sync_block();
  }
}

PAL_DEV double sync_nan() {
#if PAL_GPU
  return __longlong_as_double(0x7ff8000000000000LL);
#else
  return std::nan("");
#endif
}

struct SyncPickSmem {
  double sv[32];
  int si[32];
};

// (ii) arg-max of |corr| in scipy's 'full' order.  corr rows are in FFT order (index = lag mod n);
// 'full' index i <-> lag i - (len_ref - 1), i in [0, len_m + len_ref - 2].  Output per item:
// peak_index (first maximum, np.argmax :427), absmax = |corr[peak_index]|, win[5] =
// corr[peak_index-2 .. peak_index+2] (NaN outside the row).
template <typename T, int NT>
PAL_DEV void sync_pick_body(const T* corr, int n, long long n_rows, long long item0, int Mics, const int* ref, int len,
                            const int* lens, int* peak_index, double* absmax, double* win, char* smem) {
  SyncPickSmem* sm = reinterpret_cast<SyncPickSmem*>(smem);
  for (long long rw = simt::bid(); rw < n_rows; rw += simt::nblocks()) {
    const long long item = item0 + rw;
    const long long f = item / Mics;
    const int lm = lens ? lens[item] : len;
    const int lr = lens ? lens[f * Mics + ref[f]] : len;
    const int full = lm + lr - 1;
    const T* c = corr + rw * n;
    double bv = -1.0;
    int bi = -1;
    for (int i = simt::tid(); i < full; i += NT) {
      const int lag = i - (lr - 1);
      const double v = fabs(double(c[lag < 0 ? lag + n : lag]));
      if (v > bv) { bv = v; bi = i; }
    }
    block_argmax<false, double, NT>(bv, bi, sm->sv, sm->si);
    if (simt::tid() < 5) {
      const int i = bi - 2 + simt::tid();
      double v = sync_nan();
      if (bi >= 0 && i >= 0 && i < full) {
        const int lag = i - (lr - 1);
        v = double(c[lag < 0 ? lag + n : lag]);
      }
      win[item * 5 + simt::tid()] = v;
    }
    if (simt::tid() == 0) {
      peak_index[item] = bi;
      absmax[item] = bv;
    }
    simt::sync_block();
  }
}

// (iv) out[row][pad[row] + j] = in[row][j] for j < lens[row], zero elsewhere   (np.pad, :452-457)
template <typename T>
PAL_DEV void pad_rows_body(const T* in, long long n_rows, long long ld_in, int len, const int* lens, const int* pad,
                           T* out, long long ld_out) {
  const long long total = n_rows * ld_out;
  for (long long e = (long long)simt::bid() * simt::nthreads() + simt::tid(); e < total;
       e += (long long)simt::nblocks() * simt::nthreads()) {
    const long long row = e / ld_out;
    const long long j = e - row * ld_out - pad[row];
    const int l = lens ? lens[row] : len;
    out[e] = (j >= 0 && j < l) ? in[row * ld_in + j] : T(0);
  }
}

}  // namespace pal
