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

// ---- per-channel whitening (the fast path's form of the PHAT weighting, as on the n = 4095 path) ------------------
// |S_i conj(S_j)| = |S_i| |S_j|, so R / (|R| + 1e-10) = U_i conj(U_j) g with U = S / |S| and g = m / (m + 1e-10),
// m = |S_i| |S_j|.  Every channel is unpacked from its packed transform and normalised ONCE (M times per frame instead
// of P times), only the Hermitian half is kept, and the pair loader multiplies two unit phasors: 2 loads and one complex
// product per pair and bin instead of 4 loads, an unpack and a reciprocal square root.  The factor g is not ignored but
// BOUNDED per row: |corr_U[k] - corr[k]| <= (1/n) sum_k 1e-10 / m_k <= 1e-10 sqrt(h_i h_j) (Cauchy-Schwarz) with
// h = mean_k 1 / |S_k|^2 at the signal's true level, one float per channel.  The pick widens its near-tie margin by the
// bound and sends the row to the float64 sweep when the bound exceeds 2 eps (very quiet channels, where the reference's
// absolute 1e-10 makes the result level-dependent); a bin with S = 0 gives U = 0 (R = 0 in both forms) and h = inf.
// The rounding noise of the float32 forward transform is bounded the same way as on the n = 4095 path
// (pal_pfa4095.cuh: whiten_bin): q = min(4, sigma^2 mean_k 1 / |S_k|^2), sigma^2 = 2^-48 mean_k |S_k|^2 (Parseval), one
// more float per channel, gathered in the same pass; the pick adds kNoiseK sqrt((q_i + q_j) / n) to the margin.
// Z: packed spectra [n_packed][n]; frame-major: packed row g <-> frame g / CP, channels 2c, 2c+1 (c = g % CP).
// U: [frames * Mics][Hn], Hn = n / 2 + 1 (nullptr: statistics only, for the sweeps that keep the packed spectra);
// hq: [frames * Mics][2] = (h, q); scales: [.][2] of the channels (global rows from row_base).
template <int NT>
PAL_DEV void whiten_unpack_body(const cpxf* Z, int n, long long n_packed, int Mics, int CP, const float* scales, long long row_base,
                                long long local_row_base, cpxf* U, float* hq, char* smem) {
  float* sh = reinterpret_cast<float*>(smem);      // [4][NT / 32]
  constexpr int NW = NT / 32;
  const int Hn = n / 2 + 1;
  for (long long g = simt::bid(); g < n_packed; g += simt::nblocks()) {
    const long long f = g / CP;
    const int c = int(g - f * CP);
    const long long ra = local_row_base + f * Mics + 2 * c;            // resident channel rows of the pair
    const bool has_b = 2 * c + 1 < Mics;
    const cpxf* z = Z + g * n;
    cpxf* ua = U ? U + ra * Hn : nullptr;
    cpxf* ub = U ? U + (ra + 1) * Hn : nullptr;
    float sa = 0.f, sb = 0.f, qa = 0.f, qb = 0.f;        // sums of 1 / |S|^2 and of |S|^2
    for (int k = simt::tid(); k < Hn; k += NT) {
      const cpxf a = unpack_two_real<float>(z, n, k, false);
      const cpxf b = unpack_two_real<float>(z, n, k, true);
      const float wgt = (k == 0 || 2 * k == n) ? 1.f : 2.f;           // bins k and n - k carry the same magnitude
      const float ma = fma_(a.x, a.x, a.y * a.y), mb = fma_(b.x, b.x, b.y * b.y);
      float ia = 0.f, ib = 0.f;
#if PAL_GPU
      if (ma > 0.f) ia = rsqrtf(ma);
      if (mb > 0.f) ib = rsqrtf(mb);
#else
      if (ma > 0.f) ia = 1.f / std::sqrt(ma);
      if (mb > 0.f) ib = 1.f / std::sqrt(mb);
#endif
      if (U) {
        ua[k] = cpxf{a.x * ia, a.y * ia};
        if (has_b) ub[k] = cpxf{b.x * ib, b.y * ib};
      }
      sa += wgt * (ma > 0.f ? ia * ia : 3.0e38f);
      sb += wgt * (mb > 0.f ? ib * ib : 3.0e38f);
      qa = fma_(wgt, ma, qa);
      qb = fma_(wgt, mb, qb);
    }
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) {
      sa += simt::shfl_xor(sa, m);
      sb += simt::shfl_xor(sb, m);
      qa += simt::shfl_xor(qa, m);
      qb += simt::shfl_xor(qb, m);
    }
    if (simt::lane() == 0) {
      sh[simt::warp()] = sa; sh[NW + simt::warp()] = sb;
      sh[2 * NW + simt::warp()] = qa; sh[3 * NW + simt::warp()] = qb;
    }
    simt::sync_block();
    if (simt::tid() == 0) {
      float ta = 0.f, tb = 0.f, ua_ = 0.f, ub_ = 0.f;
      for (int w = 0; w < NW; ++w) { ta += sh[w]; tb += sh[NW + w]; ua_ += sh[2 * NW + w]; ub_ += sh[3 * NW + w]; }
      // back to the signal's own level: S_true = S_scaled * 2^e, so 1 / |S_true|^2 = 2^-2e / |S_scaled|^2 (q is level-free)
      const long long gra = row_base + f * Mics + 2 * c;
      const float ea = scales[2 * gra], eb = has_b ? scales[2 * (gra + 1)] : 0.f;
      ta /= float(n);
      tb /= float(n);
      // q = sigma^2 mean(1 / |S|^2), sigma^2 = 2^-48 mean |S|^2 (Parseval), capped at 4; NaN (0 * inf) ends as 4
      const float qa_ = 3.5527137e-15f * (ua_ / float(n)) * ta, qb_ = 3.5527137e-15f * (ub_ / float(n)) * tb;
      hq[2 * ra] = ta * ea * ea;
      hq[2 * ra + 1] = (ua_ > 0.f) ? min_(4.f, qa_) : 0.f;
      if (has_b) {
        hq[2 * ra + 2] = tb * eb * eb;
        hq[2 * ra + 3] = (ub_ > 0.f) ? min_(4.f, qb_) : 0.f;
      }
    }
    simt::sync_block();
  }
}

// inverse DFT of TWO whitened cross spectra per transform: a[k] = (R_A[k] + i R_B[k]) conj(chirp[k]) / n with
// R[k] = U_i[k] conj(U_j[k]) (k <= n/2) and R[n - k] = conj(R[k]); items A = t_off + 2t, B = A + 1 (frame-major).
struct LoadPhatU {
  BluePlan p;
  const cpxf* chirp;
  const cpxf* U;           // [resident frames * Mics][Hn]
  const int* pairs;        // [P][2]
  int Mics, P, Hn;
  long long t_off, n_items;
  const float* scales;     // [all rows][2], indexed by GLOBAL channel row (dead channels)
  long long frame_base;    // global index of resident frame 0
  struct Item {
    const cpxf *ui, *uj;
    bool present, dead;
  };
  PAL_DEV Item item(long long it) const {
    Item m;
    m.present = it < n_items;
    if (!m.present) {
      m.ui = m.uj = U;
      m.dead = true;
      return m;
    }
    const long long f = it / P;
    const int pr = int(it - f * P);
    const int mi = pairs[2 * pr], mj = pairs[2 * pr + 1];
    m.ui = U + (f * Mics + mi) * Hn;
    m.uj = U + (f * Mics + mj) * Hn;
    m.dead = scales[2 * ((frame_base + f) * Mics + mi)] == 0.f || scales[2 * ((frame_base + f) * Mics + mj)] == 0.f;
    return m;
  }
  struct Ctx {
    Item a, b;
  };
  PAL_DEV Ctx begin(long long t) const {
    const long long ia = t_off + 2 * t;
    return Ctx{item(ia), item(ia + 1)};
  }
  // branch-free on purpose: every load is unconditional (absent / dead items point at valid memory and are masked
  // afterwards), so the compiler can issue the loads of several samples back to back instead of one round trip at a time
  PAL_DEV cpxf operator()(const Ctx& c, int k) const {
    const bool in = k < p.n;
    const int kc = in ? k : 0;
    const bool up = 2 * kc > p.n;                // upper half: R[k] = conj(R[n - k])
    const int kk = up ? p.n - kc : kc;
    const cpxf a1 = c.a.ui[kk], b1 = c.a.uj[kk], a2 = c.b.ui[kk], b2 = c.b.uj[kk], w = chirp[kc];
    const float ma = (in && !c.a.dead) ? 1.f / float(p.n) : 0.f;
    const float mb = (in && c.b.present && !c.b.dead) ? 1.f / float(p.n) : 0.f;
    const float sg = up ? -1.f : 1.f;
    const float rax = fma_(a1.x, b1.x, a1.y * b1.y) * ma, ray = fma_(a1.y, b1.x, -(a1.x * b1.y)) * (ma * sg);     // a conj(b)
    const float rbx = fma_(a2.x, b2.x, a2.y * b2.y) * mb, rby = fma_(a2.y, b2.x, -(a2.x * b2.y)) * (mb * sg);
    const float xr = rax - rby, xi = ray + rbx;
    return cpxf{fma_(xr, w.x, xi * w.y), fma_(xi, w.x, -(xr * w.y))};          // (R_A + i R_B) conj(chirp)
  }
};

// packed inverse transform -> window samples + per-tile row maxima (see the header comment).  Protocol of a storer of
// pal_fft2.cuh: begin(t) per work unit, operator() per sample, finish() once per unit by the whole block.
template <class Src> struct StoreWin2T {
  BluePlan p;
  const cpxf* chirp;
  float* win;              // [rows of this launch][g.wstride]
  float* pmax;             // [rows of this launch][tiles]
  long long n_rows;        // rows of this launch (the last transform may own a single row)
  Src src;                 // what was transformed: a dead item's row is exact zeros (the reference's R = 0)
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

using StoreWin2 = StoreWin2T<LoadPhat2<float>>;
using StoreWinU = StoreWin2T<LoadPhatU>;

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
// hq != nullptr: the rows come from whitened spectra (LoadPhatU); row r is item first_item + r of the resident frames,
// hq [resident frames * Mics] holds the per-channel whitening bound (whiten_unpack_body)
struct WhitenRef {
  const float* hq;
  const int* pairs;
  int Mics, P;
  long long first_item;
  bool whitened = true;      // false: the rows come from the raw spectra (LoadPhat2), only the rounding-noise term applies
};
// the chirp-z forward transforms (two FFTs of 2-4 n points and three chirp products) are about twice as noisy as the
// prime-factor transform of the n = 4095 path: rms error / model 1.5 .. 3.5, largest sample error <= 11.3 deviations
constexpr float kWinNoiseK = 40.f;
// what the float32 sweep neglects for the pair of item `it` (normalised correlation units): factor g of the whitened
// spectra + kWinNoiseK standard deviations of the forward transforms' rounding noise
PAL_DEV float pair_bound(const WhitenRef& wr, long long it, int n) {
  const long long f = it / wr.P;
  const int pr = int(it - f * wr.P);
  const float* a = wr.hq + 2 * (f * wr.Mics + wr.pairs[2 * pr]);
  const float* b = wr.hq + 2 * (f * wr.Mics + wr.pairs[2 * pr + 1]);
  const float hb = wr.whitened ? 1e-10f * sqrt_(a[0] * b[0]) : 0.f;
  return hb + kWinNoiseK * sqrt_((a[1] + b[1]) / float(n));
}
template <int NT>
PAL_DEV void win_pick_rows_body(const float* win, const float* pmax, int tiles, long long n_rows, WinGeom g, long long item0,
                                int* k_idx, int* k_count, float* peak, float* gmax, unsigned* flags, unsigned extra_flag,
                                WhitenRef wr = WhitenRef{nullptr, nullptr, 0, 0, 0}) {
  const int lane = simt::lane();
  const int lo = g.lo, hi = g.hi, dist = g.dist;
  const int g_lo = (lo + 3) & ~3, g_hi = (hi + 1) & ~3;      // window split into 16-byte groups + <= 3 + 3 edge samples
  for (long long row = (long long)simt::bid() * (NT / 32) + simt::warp(); row < n_rows; row += (long long)simt::nblocks() * (NT / 32)) {
    const float* c = win + row * g.wstride - g.wlo;            // c[k] valid for wlo <= k <= whi
    float gm = kWinNegBig;
    for (int t = lane; t < tiles; t += 32) gm = max_(gm, pmax[row * tiles + t]);
    gm = wmax_f(gm);
    float eps = g.eps;
    bool quiet = false;
    if (wr.hq) {
      // small: widens the near-tie margin.  Not small against the margin and the row's maximum (or inf / NaN): the
      // VALUES may be off by a third of the 1e-4 tolerance (the bound is >= 3.5 x the largest error seen) -> float64 sweep
      const float hb = pair_bound(wr, wr.first_item + row, g.n);
      quiet = !(hb <= 2.f * g.eps + 1e-4f * max_(gm, 0.f));
      if (!quiet) eps += hb;
    }
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
    unsigned fl = extra_flag | (quiet ? PAL_FLAG_NEAR_TIE : 0u);
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
