// pal_peakpick.cuh -- the reference's TDOA peak selection, exactly, for one correlation row
// handled by one thread block.
//
// Replaces utils.py:140-181 (get_time_delays_phat after phat_correlation) including the
// behaviour of scipy.signal.find_peaks(height=, distance=) that it calls (utils.py:152,156,167):
//   * local maxima with plateau -> floor-midpoint, end samples never peaks
//   * height filter inclusive (>=), then the greedy highest-first distance rule
//   * threshold = mult*median|c| ('adaptive': mult*(mean|c| + std|c|)), alternative mean|c|
//   * window filter |k-(n2-1)| <= win_half, fallbacks down to the unbounded first argmax
// The control flow is the reduced form proved equal to the reference in
// oracle/pal_oracle.py::tdoa_pick_restated (tests/test_oracle.py::test_fuzz_control_flow).
#pragma once
#include "pal_simt.h"
#include "../../include/pal_b200.h"

namespace pal {

// per-row flag bits (PAL_FLAG_*) come from the public header include/pal_b200.h

template <typename T> PAL_DEV T warp_sum(T v) {
#pragma unroll
  for (int m = 16; m >= 1; m >>= 1) v += simt::shfl_xor(v, m);
  return v;
}

// Deterministic block sum: warp shuffles, then every thread adds the per-warp partials in
// warp order.  `scratch` holds NT/32 elements of T.
template <typename T, int NT> PAL_DEV T block_sum(T v, T* scratch) {
  v = warp_sum(v);
  if (simt::lane() == 0) scratch[simt::warp()] = v;
  simt::sync_block();
  T r = T(0);
#pragma unroll
  for (int w = 0; w < NT / 32; ++w) r += scratch[w];
  simt::sync_block();
  return r;
}

// (value, index) maximum.  TIE_HIGH: equal values -> larger index wins (scipy's priority
// order); otherwise the smaller index wins (np.argmax = first maximum).  idx < 0 = "none".
template <bool TIE_HIGH, typename T> PAL_DEV bool vi_better(T v, int i, T bv, int bi) {
  if (i < 0) return false;
  if (bi < 0) return true;
  if (v > bv) return true;
  if (v < bv) return false;
  return TIE_HIGH ? (i > bi) : (i < bi);
}
template <bool TIE_HIGH, typename T> PAL_DEV void warp_argmax(T& v, int& i) {
#pragma unroll
  for (int m = 16; m >= 1; m >>= 1) {
    T ov = simt::shfl_xor(v, m);
    int oi = simt::shfl_xor(i, m);
    if (vi_better<TIE_HIGH>(ov, oi, v, i)) { v = ov; i = oi; }
  }
}
template <bool TIE_HIGH, typename T, int NT>
PAL_DEV void block_argmax(T& v, int& i, T* sv, int* si) {
  warp_argmax<TIE_HIGH>(v, i);
  if (simt::lane() == 0) { sv[simt::warp()] = v; si[simt::warp()] = i; }
  simt::sync_block();
  v = sv[0]; i = si[0];
#pragma unroll
  for (int w = 1; w < NT / 32; ++w)
    if (vi_better<TIE_HIGH>(sv[w], si[w], v, i)) { v = sv[w]; i = si[w]; }
  simt::sync_block();
}

template <typename T> struct KeyBits;
template <> struct KeyBits<float> {
  using U = unsigned;
  static constexpr int TOP = 30;  // |x| has sign bit 0
  static PAL_DEV U of(float x) {
#if PAL_GPU
    return __float_as_uint(x);
#else
    U u; std::memcpy(&u, &x, 4); return u;
#endif
  }
  static PAL_DEV float from(U u) {
#if PAL_GPU
    return __uint_as_float(u);
#else
    float x; std::memcpy(&x, &u, 4); return x;
#endif
  }
};
template <> struct KeyBits<double> {
  using U = unsigned long long;
  static constexpr int TOP = 62;
  static PAL_DEV U of(double x) {
#if PAL_GPU
    return (U)__double_as_longlong(x);
#else
    U u; std::memcpy(&u, &x, 8); return u;
#endif
  }
  static PAL_DEV double from(U u) {
#if PAL_GPU
    return __longlong_as_double((long long)u);
#else
    double x; std::memcpy(&x, &u, 8); return x;
#endif
  }
};

// rank-th smallest (0-based) of |c[0..n)| by bitwise binary search on the IEEE pattern
// (non-negative floats order like unsigned integers).  Exact, no sort, no extra storage.
template <typename T, int NT> PAL_DEV T select_abs(const T* c, int n, int rank, int* iscratch) {
  using K = KeyBits<T>;
  typename K::U res = 0;
  for (int b = K::TOP; b >= 0; --b) {
    const typename K::U cand = res | (typename K::U(1) << b);
    int cnt = 0;
    for (int k = simt::tid(); k < n; k += NT) cnt += (K::of(abs_(c[k])) < cand) ? 1 : 0;
    cnt = block_sum<int, NT>(cnt, iscratch);
    if (cnt <= rank) res = cand;
  }
  return K::from(res);
}

struct PickResult {
  int count;        // number of indices written to out_k (<= num_peaks)
  unsigned flags;
  double sum_abs;   // sum |c| over the row (pass A), for callers that audit the decision
};

struct PickScratch {  // lives in shared memory
  double dsum[32];
  int isum[32];
  int iarg[32];
  int bcast[4];
};

// threshold_multiplier * median|c|  ('adaptive': * (mean|c| + population std|c|)), utils.py:144-149
template <typename T, int NT>
PAL_DEV T pick_threshold(const T* c, int n, int method, T mult, T mean_abs, PickScratch* ps) {
  const int tid = simt::tid();
  T* tsum = reinterpret_cast<T*>(ps->dsum);
  T thr;
  if (method == 1) {
    T ss = T(0);
    for (int k = tid; k < n; k += NT) {
      const T d = abs_(c[k]) - mean_abs;
      ss = fma_(d, d, ss);
    }
    ss = block_sum<T, NT>(ss, tsum);
    thr = mult * (mean_abs + sqrt_(ss / T(n)));
  } else {
    T med;
    if (n & 1) {
      med = select_abs<T, NT>(c, n, (n - 1) / 2, ps->isum);
    } else {
      const T lo = select_abs<T, NT>(c, n, n / 2 - 1, ps->isum);
      int cnt = 0;
      T nxt = T(0);
      int ni = -1;
      for (int k = tid; k < n; k += NT) {
        const T a = abs_(c[k]);
        if (a <= lo) ++cnt;
        else if (vi_better<false>(-a, k, nxt, ni)) { nxt = -a; ni = k; }   // min of those > lo
      }
      cnt = block_sum<int, NT>(cnt, ps->isum);
      block_argmax<false, T, NT>(nxt, ni, tsum, ps->iarg);
      const T hi = (cnt >= n / 2 + 1 || ni < 0) ? lo : -nxt;
      med = (lo + hi) * T(0.5);
    }
    thr = mult * med;
  }
  return thr;
}

// One block, NT threads.  c: row of n values (shared or global memory).  pkmap: n bytes of
// scratch (shared or global).  Returns (in every thread) the result; out_k written by thread 0.
template <typename T, int NT>
PAL_DEV PickResult peakpick_row(const T* c, int n, int c0, int win_half, int dist, int method, T mult,
                                int num_peaks, unsigned char* pkmap, PickScratch* ps, int* out_k,
                                T* out_gmax, T* out_peak) {
  const int tid = simt::tid();
  T* tsum = reinterpret_cast<T*>(ps->dsum);
  unsigned flags = 0;

  // ---- pass A: mean|c|, global max (first index), zero the peak map --------------------
  T s_abs = T(0);
  T gm = T(0);
  int gi = -1;
  for (int k = tid; k < n; k += NT) {
    const T v = c[k];
    s_abs += abs_(v);
    if (vi_better<false>(v, k, gm, gi)) { gm = v; gi = k; }
    pkmap[k] = 0;
  }
  s_abs = block_sum<T, NT>(s_abs, tsum);
  block_argmax<false, T, NT>(gm, gi, tsum, ps->iarg);
  const T mean_abs = s_abs / T(n);

  // ---- local maxima with plateaus (scipy _local_maxima_1d) -> pkmap, highest peak ------
  T gpk = T(0);
  int gpk_i = -1;
  for (int i = tid + 1; i < n - 1; i += NT) {
    const T v = c[i];
    if (c[i - 1] < v) {
      int j = i + 1;
      while (j < n - 1 && c[j] == v) ++j;
      if (c[j] < v) {
        const int mid = (i + j - 1) >> 1;
        pkmap[mid] = 1;
        if (vi_better<true>(v, mid, gpk, gpk_i)) { gpk = v; gpk_i = mid; }
      }
    }
  }
  block_argmax<true, T, NT>(gpk, gpk_i, tsum, ps->iarg);   // also a barrier: pkmap complete

  if (tid == 0) *out_gmax = gm;
  // The median costs 31 (63 in float64) counting sweeps, and for num_peaks == 1 the answer does
  // not depend on it whenever the best surviving in-window peak reaches mean|c| (it is then
  // accepted either directly or through the alternative threshold, utils.py:155,166).  So the
  // threshold is evaluated lazily; every condition below is block-uniform.
  bool have_thr = false;
  T thr = T(0), thr_eff = T(0);
  if (gpk_i >= 0 && (num_peaks != 1 || gpk < mean_abs)) {
    thr = pick_threshold<T, NT>(c, n, method, mult, mean_abs, ps);
    have_thr = true;
  }
  if (gpk_i < 0 || (have_thr && gpk < thr && gpk < mean_abs)) {          // utils.py:153-160
    if (tid == 0) { out_k[0] = gi; *out_peak = gm; }
    PickResult r{1, flags | PAL_FLAG_FALLBACK_ARGMAX, double(s_abs)};
    return r;
  }
  if (have_thr) {
    thr_eff = thr;
    if (gpk < thr) { thr_eff = mean_abs; flags |= PAL_FLAG_ALT_THRESHOLD; }
  }

  int lo = 1, hi = n - 2;
  if (win_half >= 0) {
    lo = (c0 - win_half > 1) ? c0 - win_half : 1;
    hi = (c0 + win_half < n - 2) ? c0 + win_half : n - 2;
  } else if (win_half < -1) {
    lo = 1; hi = 0;   // empty window
  }

  // ---- enumerate in-window peaks by descending priority; keep the survivors --------------
  int count = 0;
  T prev_v = T(0);
  int prev_i = -1;          // -1: no upper bound yet
  bool relaxed = false;     // threshold already lowered to mean|c| (utils.py:164-167)
  for (;;) {
    T bv = T(0);
    int bi = -1;
    for (int k = lo + tid; k <= hi; k += NT) {
      if (!pkmap[k]) continue;
      const T v = c[k];
      if (prev_i >= 0 && !vi_better<true>(prev_v, prev_i, v, k)) continue;   // need prio(k) < prio(prev)
      if (vi_better<true>(v, k, bv, bi)) { bv = v; bi = k; }
    }
    block_argmax<true, T, NT>(bv, bi, tsum, ps->iarg);
    bool stop = false;
    if (bi >= 0 && !have_thr && bv < mean_abs) {     // lazy threshold needed after all
      thr = pick_threshold<T, NT>(c, n, method, mult, mean_abs, ps);
      have_thr = true;
      thr_eff = thr;
      if (gpk < thr) { thr_eff = mean_abs; flags |= PAL_FLAG_ALT_THRESHOLD; }
    }
    if (bi < 0) {
      stop = true;
    } else if (have_thr && bv < thr_eff) {
      if (count == 0 && win_half != -1 && !relaxed && bv >= mean_abs) {
        relaxed = true;                     // retry with the alternative threshold
        thr_eff = mean_abs;
        flags |= PAL_FLAG_ALT_THRESHOLD;
      } else {
        stop = true;
      }
    }
    if (stop) break;
    // survival of peak bi under the greedy distance rule: alive iff no ALIVE peak of higher
    // priority lies closer than `dist`.  Depth-first over strictly increasing priorities.
    if (tid == 0) {
      constexpr int MAXD = 48;
      int st_p[MAXD], st_o[MAXD];
      int depth = 0;
      st_p[0] = bi;
      st_o[0] = -(dist - 1);
      int verdict = -1;  // alive status of the frame just popped
      bool overflow = false;
      while (depth >= 0) {
        const int p = st_p[depth];
        if (verdict == 1) {          // a killer of p is alive -> p is dead
          verdict = 0;
          --depth;
          continue;
        }
        verdict = -1;
        bool pushed = false;
        int o = st_o[depth];
        for (; o <= dist - 1; ++o) {
          const int q = p + o;
          if (o == 0 || q < 1 || q > n - 2 || !pkmap[q]) continue;
          if (!vi_better<true>(c[q], q, c[p], p)) continue;
          st_o[depth] = o + 1;
          if (depth + 1 >= MAXD) { overflow = true; break; }
          ++depth;
          st_p[depth] = q;
          st_o[depth] = -(dist - 1);
          pushed = true;
          break;
        }
        if (overflow) break;
        if (pushed) continue;
        verdict = 1;                 // scan exhausted: p is alive
        --depth;
      }
      // after the loop `verdict` is the status of the root frame
      ps->bcast[0] = overflow ? 2 : verdict;
    }
    simt::sync_block();
    const int alive = ps->bcast[0];
    simt::sync_block();
    if (alive == 2) flags |= PAL_FLAG_STACK_OVERFLOW;
    if (alive >= 1) {
      if (tid == 0) {
        out_k[count] = bi;
        if (count == 0) *out_peak = bv;
      }
      ++count;
      if (count >= num_peaks) break;
    }
    prev_v = bv;
    prev_i = bi;
  }
  if (count == 0) {                                          // utils.py:168-172 (unbounded argmax)
    if (tid == 0) { out_k[0] = gi; *out_peak = gm; }
    count = 1;
    flags |= PAL_FLAG_FALLBACK_ARGMAX;
  }
  PickResult r{count, flags, double(s_abs)};
  return r;
}

}  // namespace pal
