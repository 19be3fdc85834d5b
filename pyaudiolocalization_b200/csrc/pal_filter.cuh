// pal_filter.cuh -- zero-phase IIR filtering of many channels at once: scipy.signal.filtfilt(b, a, x)
// with its defaults (method "pad", odd extension of 3 * ntaps samples, lfilter_zi initial state),
// which is what signal_processing.noise_reduction(..., 'butterworth') applies to every channel
// between the two stages of localize_sound_source (signal_processing.py:124-128, main.py:191).
//
// The recursion is sequential along time but independent across channels: one THREAD per channel
// (row), float64, the direct-form-II-transposed update of scipy's lfilter in exactly its evaluation
// order and without FMA contraction, so the result is bit-identical to scipy on x86:
//     y    = z[0] + b[0] x
//     z[i] = (z[i+1] + x b[i+1]) - y a[i+1]          i = 0 .. ntaps-3
//     z[ntaps-2] = x b[ntaps-1] - y a[ntaps-1]
// A warp handles 32 rows; rows are read / written through padded shared-memory tiles so that global
// accesses stay coalesced, and the forward-pass output is kept in a workspace laid out [k][32 rows].
#pragma once
#include "pal_simt.h"

namespace pal {

constexpr int kFiltMaxTaps = 16;

struct FiltParams {
  int ntaps;    // max(len(a), len(b)), both zero-padded to it; a[0] == 1
  int padlen;   // 3 * ntaps for scipy's default
  double b[kFiltMaxTaps], a[kFiltMaxTaps], zi[kFiltMaxTaps];   // zi: lfilter_zi(b, a), ntaps-1 values
};

#if PAL_GPU
PAL_DEV double fmul_rn(double a, double b) { return __dmul_rn(a, b); }
PAL_DEV double fadd_rn(double a, double b) { return __dadd_rn(a, b); }
#else
inline double fmul_rn(double a, double b) { volatile double r = a * b; return r; }
inline double fadd_rn(double a, double b) { volatile double r = a + b; return r; }
#endif

struct FiltState {
  double z[kFiltMaxTaps];
};
PAL_DEV double filt_step(const FiltParams& fp, FiltState& s, double x) {
  const double y = fadd_rn(s.z[0], fmul_rn(fp.b[0], x));
#pragma unroll
  for (int i = 0; i < kFiltMaxTaps - 1; ++i) {      // compile-time indices only: z stays in registers
    if (i < fp.ntaps - 2) s.z[i] = fadd_rn(fadd_rn(s.z[i + 1], fmul_rn(x, fp.b[i + 1])), -fmul_rn(y, fp.a[i + 1]));
    else if (i == fp.ntaps - 2) s.z[i] = fadd_rn(fmul_rn(x, fp.b[i + 1]), -fmul_rn(y, fp.a[i + 1]));
  }
  return y;
}

// One warp = 32 rows.  tile: [32][33] doubles of shared memory per warp.
template <typename TIO> struct RowTile {
  double* t;           // [32][33]
  const TIO* x;        // rows of n samples
  TIO* y;
  long long row0, n_rows;
  int n;
  int cur;             // index of the 32-sample block held in the tile (-1: none)
  bool dirty;
  PAL_DEV void load(int blk) {
    const int lane = simt::lane();
    const int c = blk * 32 + lane;
    for (int r = 0; r < 32; ++r)
      t[r * 33 + lane] = (row0 + r < n_rows && c < n) ? double(x[(row0 + r) * n + c]) : 0.0;
    simt::sync_warp();
    cur = blk;
  }
  PAL_DEV void flush() {
    const int lane = simt::lane();
    simt::sync_warp();
    const int c = cur * 32 + lane;
    for (int r = 0; r < 32; ++r)
      if (row0 + r < n_rows && c < n) y[(row0 + r) * n + c] = TIO(t[r * 33 + lane]);
    simt::sync_warp();
    dirty = false;
  }
  PAL_DEV double get(int i) {      // sample i of this lane's row (all lanes ask for the same i)
    if ((i >> 5) != cur) { simt::sync_warp(); load(i >> 5); }
    return t[simt::lane() * 33 + (i & 31)];
  }
  PAL_DEV void put(int i, double v) {
    if ((i >> 5) != cur) {
      if (dirty) flush();
      simt::sync_warp();
      cur = i >> 5;
    }
    t[simt::lane() * 33 + (i & 31)] = v;
    dirty = true;
  }
};

// work: [groups][n + 2 padlen][32] doubles, group = 32 consecutive rows
template <typename TIO, int NT>
PAL_DEV void filtfilt_body(const TIO* x, long long n_rows, int n, FiltParams fp, double* work, TIO* y, char* smem_raw) {
  const int lane = simt::lane();
  const int p = fp.padlen;
  const int L = n + 2 * p;
  double* tile = reinterpret_cast<double*>(smem_raw) + size_t(simt::warp()) * 32 * 33;
  const long long groups = (n_rows + 31) / 32;
  const long long warps_total = (long long)simt::nblocks() * (NT / 32);
  for (long long g = (long long)simt::bid() * (NT / 32) + simt::warp(); g < groups; g += warps_total) {
    RowTile<TIO> in{tile, x, nullptr, g * 32, n_rows, n, -1, false};
    double* w = work + size_t(g) * L * 32 + lane;
    const double x0 = in.get(0);
    const double xl = in.get(n - 1);
    // ---- forward pass over the odd extension (scipy odd_ext): 2 x[0] - x[p..1], x, 2 x[n-1] - x[n-2..n-1-p]
    FiltState s;
    const double e0 = fadd_rn(fmul_rn(2.0, x0), -in.get(p));
#pragma unroll
    for (int i = 0; i < kFiltMaxTaps; ++i) s.z[i] = (i < fp.ntaps - 1) ? fmul_rn(fp.zi[i], e0) : 0.0;
    for (int k = 0; k < L; ++k) {
      double v;
      if (k < p) v = fadd_rn(fmul_rn(2.0, x0), -in.get(p - k));
      else if (k < p + n) v = in.get(k - p);
      else v = fadd_rn(fmul_rn(2.0, xl), -in.get(n - 2 - (k - p - n)));
      w[size_t(k) * 32] = filt_step(fp, s, v);
    }
    simt::sync_warp();
    // ---- backward pass over the reversed forward output; keep y[p .. p+n)
    RowTile<TIO> out{tile, nullptr, y, g * 32, n_rows, n, -1, false};
    const double y0 = w[size_t(L - 1) * 32];
#pragma unroll
    for (int i = 0; i < kFiltMaxTaps; ++i) s.z[i] = (i < fp.ntaps - 1) ? fmul_rn(fp.zi[i], y0) : 0.0;
    for (int m = 0; m < L; ++m) {
      const double t = filt_step(fp, s, w[size_t(L - 1 - m) * 32]);
      const int i = L - 1 - p - m;            // output sample index
      if (i >= 0 && i < n) out.put(i, t);
    }
    if (out.dirty) out.flush();
    simt::sync_warp();
  }
}

}  // namespace pal
