// pal_winpick.cuh -- the TDOA pick of the arbitrary-length path in its reduced ("fast") form: only the samples a
// decision for ONE peak can depend on ever reach global memory.
//
// utils.py:140-181 with num_peaks = 1 (what main.py:204 asks for) picks the highest find_peaks() survivor inside the
// window |lag| <= max_expected_delay.  Whatever the threshold method, that is the largest sample of the window whenever
//   (a) it is a strict local maximum and clearly above mean|c| (the threshold of the reference's retry branch,
//       utils.py:155/166; mean|c| <= rms(c) <= 1/sqrt(n) by Parseval because |R| <= 1 after PHAT weighting),
//   (b) no other sample of the window comes within eps of it, and
//   (c) no sample within `distance` of it -- inside the window or not -- comes within eps of it (find_peaks' distance
//       rule could remove it, directly or through a chain).
// Everything else (near ties, plateaus, chains, a winner near the thresholds, no peak in the window at all: the
// reference's unbounded-argmax fallbacks) is flagged and re-evaluated by the float64 sweep with the complete
// find_peaks emulation (pal_peakpick.cuh).  The same reduction drives the fused n = 4095 kernel (pal_pfa4095.cuh:
// fast_pick_row); here the row comes from the inverse column pass of the convolution engine, whose storer keeps
//   * the window plus a margin of `distance` + 1 samples on both sides (StoreWin2::win), and
//   * the row maximum, as one partial per column tile (main.py:223 needs max(corr) of the WHOLE row),
// instead of writing all n samples and reading them back: at n = 88199, W = 2205 that is 5 % of the row.
#pragma once
#include "pal_bluestein.cuh"

namespace pal {

constexpr float kWinNegBig = -3.0e38f;

struct WinGeom {
  int n, c0;        // row length, index of lag 0 (n2 - 1)
  int lo, hi;       // candidate positions (window clipped to the interior 1 .. n-2); hi < lo: empty window
  int wlo, whi;     // samples kept: [wlo, whi], wlo a multiple of 4
  int wstride;      // floats per kept row (multiple of 4)
  int dist;
  float eps, mean_bound;
};
PAL_HD WinGeom make_win_geom(int n, int c0, int win_half, int dist, float eps) {
  WinGeom g;
  g.n = n;
  g.c0 = c0;
  g.dist = dist;
  g.eps = eps;
  g.lo = 1;
  g.hi = n - 2;
  if (win_half >= 0) {
    g.lo = (c0 - win_half > 1) ? c0 - win_half : 1;
    g.hi = (c0 + win_half < n - 2) ? c0 + win_half : n - 2;
  } else if (win_half < -1) {
    g.lo = 1;
    g.hi = 0;
  }
  if (g.hi >= g.lo) {
    const int a = g.lo - dist - 1, b = g.hi + dist + 1;
    g.wlo = (a > 0 ? a : 0) & ~3;
    g.whi = b < n - 1 ? b : n - 1;
  } else {
    g.wlo = 0;
    g.whi = 3;
  }
  g.wstride = (g.whi - g.wlo + 1 + 3) & ~3;
  // mean|c| <= 1/sqrt(n); a hair above it so that float32 rounding of the bound itself cannot matter
  g.mean_bound = 1.001f / sqrtf(float(n));
  return g;
}

// packed inverse transform -> window samples + per-tile row maxima (see the header comment).  Protocol of a storer of
// pal_fft2.cuh: begin(t) per work unit, operator() per sample, finish() once per unit by the whole block.
struct StoreWin2 {
  BluePlan p;
  const cpxf* chirp;
  float* win;              // [rows of this launch][g.wstride]
  float* pmax;             // [rows of this launch][tiles]
  long long n_rows;        // rows of this launch (the last transform may own a single row)
  LoadPhat2<float> src;    // what was transformed: a dead item's row is exact zeros (the reference's R = 0)
  WinGeom g;
  int tiles;
  struct Ctx {
    float *ra, *rb;        // rb == nullptr: no second row
    bool dead_a, dead_b;
    float ma, mb;
  };
  PAL_DEV Ctx begin(long long t) const {
    const long long ia = src.t_off + 2 * t;
    Ctx c;
    c.ra = win + (2 * t) * g.wstride;
    c.rb = (2 * t + 1 < n_rows) ? c.ra + g.wstride : nullptr;
    c.dead_a = src.item(ia).dead;
    c.dead_b = src.item(ia + 1).dead;
    c.ma = c.mb = kWinNegBig;
    return c;
  }
  PAL_DEV void operator()(Ctx& c, int k, cpxf y) const {
    if (k < p.n) {
      const cpxf w = chirp[k];
      const float va = c.dead_a ? 0.f : fma_(y.x, w.x, y.y * w.y);        // Re(y conj(w)) = corr_A[k]
      const float vb = c.dead_b ? 0.f : fma_(y.y, w.x, -(y.x * w.y));     // Im(y conj(w)) = corr_B[k]
      c.ma = max_(c.ma, va);
      c.mb = max_(c.mb, vb);
      if (k >= g.wlo && k <= g.whi) {
        c.ra[k - g.wlo] = va;
        if (c.rb) c.rb[k - g.wlo] = vb;
      }
    }
  }
  // block-wide maxima of the unit -> pmax[row][tile]; `scratch` holds 2 floats per warp
  PAL_DEV void finish(const Ctx& c, long long t, int tile, float* scratch) const {
    float a = c.ma, b = c.mb;
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) {
      a = max_(a, simt::shfl_xor(a, m));
      b = max_(b, simt::shfl_xor(b, m));
    }
    if (simt::lane() == 0) {
      scratch[2 * simt::warp()] = a;
      scratch[2 * simt::warp() + 1] = b;
    }
    simt::sync_block();
    if (simt::tid() == 0) {
      const int nw = simt::nthreads() / 32;
      for (int w = 1; w < nw; ++w) {
        a = max_(a, scratch[2 * w]);
        b = max_(b, scratch[2 * w + 1]);
      }
      pmax[(2 * t) * tiles + tile] = a;
      if (c.rb) pmax[(2 * t + 1) * tiles + tile] = b;
    }
    simt::sync_block();
  }
};

PAL_DEV float wmax_f(float v) {
#pragma unroll
  for (int m = 16; m >= 1; m >>= 1) v = max_(v, simt::shfl_xor(v, m));
  return v;
}
PAL_DEV int wmax_i(int v) {
#pragma unroll
  for (int m = 16; m >= 1; m >>= 1) {
    const int o = simt::shfl_xor(v, m);
    v = o > v ? o : v;
  }
  return v;
}

// One warp per row.  rows: win[row][wstride], pmax[row][tiles]; results of row r go to item item0 + r.
template <int NT>
PAL_DEV void win_pick_rows_body(const float* win, const float* pmax, int tiles, long long n_rows, WinGeom g, long long item0,
                                int* k_idx, int* k_count, float* peak, float* gmax, unsigned* flags, unsigned extra_flag) {
  const int lane = simt::lane();
  const int lo = g.lo, hi = g.hi, dist = g.dist;
  const float eps = g.eps;
  const int g_lo = (lo + 3) & ~3, g_hi = (hi + 1) & ~3;      // window split into 16-byte groups + <= 3 + 3 edge samples
  for (long long row = (long long)simt::bid() * (NT / 32) + simt::warp(); row < n_rows; row += (long long)simt::nblocks() * (NT / 32)) {
    const float* c = win + row * g.wstride - g.wlo;            // c[k] valid for wlo <= k <= whi
    float gm = kWinNegBig;
    for (int t = lane; t < tiles; t += 32) gm = max_(gm, pmax[row * tiles + t]);
    gm = wmax_f(gm);
    float b1 = kWinNegBig, b2 = kWinNegBig;
    int ksel = -1;
    if (hi >= lo) {
      int gsel = -1;
      for (int q = g_lo + 4 * lane; q < g_hi; q += 128) {
        const float4 v = *reinterpret_cast<const float4*>(c + q);
        const float m01 = max_(v.x, v.y), n01 = min_(v.x, v.y);
        const float m23 = max_(v.z, v.w), n23 = min_(v.z, v.w);
        const float m = max_(m01, m23);
        const float s4 = max_(min_(m01, m23), max_(n01, n23));       // second largest of the four
        b2 = max_(b2, max_(min_(b1, m), s4));
        if (m > b1) gsel = q;
        b1 = max_(b1, m);
      }
      if (gsel >= 0) {
        const float4 v = *reinterpret_cast<const float4*>(c + gsel);
        ksel = gsel + ((v.w == b1) ? 3 : (v.z == b1) ? 2 : (v.y == b1) ? 1 : 0);
      }
      if (lane < 6) {
        const int k = (lane < 3) ? lo + lane : g_hi + (lane - 3);
        const bool ok = (lane < 3) ? (k < g_lo && k <= hi) : (k <= hi && k >= lo && g_hi >= g_lo);
        if (ok) {
          const float v = c[k];
          b2 = max_(b2, min_(b1, v));
          if (v > b1) ksel = k;
          b1 = max_(b1, v);
        }
      }
    }
    const float bv = wmax_f(b1);
    const int bi = wmax_i((b1 == bv) ? ksel : -1);      // equal maxima: the later one (it is flagged as a tie anyway)
    const float cand2 = wmax_f((ksel == bi) ? b2 : b1);
    unsigned fl = extra_flag;
    int kbest = 0;
    float hbest = 0.f;
    if (bi >= 0) {
      kbest = bi;
      hbest = bv;
      const bool strict = c[bi - 1] < bv && bv > c[bi + 1];
      if (!strict) fl |= PAL_FLAG_PLATEAU;                           // window edge / plateau: the exact sweep sorts it out
      // mean|c| <= (gm + sqrt((n-1) (1 - gm^2))) / n: sum c^2 <= 1 (Parseval, |R| <= 1) and one sample equals gm, so the
      // other n-1 samples share at most 1 - gm^2 (Cauchy-Schwarz).  With a dominant peak outside the window -- the
      // normal case: the reference centres its window on IFFT index n2-1, not on lag 0 -- this is far below 1/sqrt(n).
      const float gp = gm > 0.f ? (gm < 1.f ? gm : 1.f) : 0.f;
      const float mb = min_(g.mean_bound, 1.001f * (gp + sqrt_(float(g.n - 1) * (1.f - gp * gp))) / float(g.n));
      if (!(bv >= mb + eps)) fl |= PAL_FLAG_NEAR_TIE;                // the median / mean thresholds decide
      if (cand2 >= bv - eps) fl |= PAL_FLAG_NEAR_TIE;
      bool hit = false;
      for (int o = -dist + lane; o <= dist; o += 32) {
        const int q = bi + o;
        if (o != 0 && (q < lo || q > hi) && q >= 0 && q < g.n && c[q] >= bv - eps) hit = true;
      }
      if (simt::ballot(hit)) fl |= PAL_FLAG_CHAIN;
    } else {
      fl |= PAL_FLAG_NEAR_TIE | PAL_FLAG_FALLBACK_ARGMAX;            // empty window: the reference's fallbacks decide
    }
    if (fl & (PAL_FLAG_PLATEAU | PAL_FLAG_CHAIN)) fl |= PAL_FLAG_NEAR_TIE;     // one bit selects the rows of the exact sweep
    if (lane == 0) {
      const long long item = item0 + row;
      k_idx[item] = kbest;
      if (k_count) k_count[item] = 1;
      peak[item] = hbest;
      gmax[item] = gm;
      flags[item] = fl;
    }
  }
}

}  // namespace pal
