// pal_pfa4095.cuh -- exact length-4095 GCC-PHAT kernels (2048-sample frames, n = n1+n2-1 = 4095).
//
//   fwd4095_body        utils.py:114-115  np.fft.fft(sig, n) of every channel (two real channels
//                                         per complex transform), spectra stored in the
//                                         Hermitian-half "[q][r]" layout the pair kernels read
//   pair4095_fast_body  utils.py:116-118 + 140-181 + main.py:223   one WARP per mic pair:
//                                         cross-spectrum, PHAT weighting, inverse DFT, peak pick,
//                                         all in registers / shared memory; R never touches HBM
//   pair4095_exact_body same semantics in one BLOCK per pair, templated on the scalar type;
//                                         in float64 it is the exact re-evaluation of the rows the
//                                         fast kernel flags (near ties below the fp32 error bound)
//
// 4095 = 5*7*9*13 = 63*65: the transform is a Good-Thomas prime-factor DFT -- no twiddle
// multiplications between stages (pal_dft_small.h).
#pragma once
#ifndef PAL_STAGGER
#define PAL_STAGGER 6000   // cycles, see pair4095_fast_body
#endif
#ifndef PAL_PREWHITEN
#define PAL_PREWHITEN 1    // 1: fwd4095 stores unit phasors S/|S| and the pair kernels only multiply (see whiten_bin)
#endif
#include "pal_dft_small.h"
#include "pal_peakpick.cuh"

namespace pal {

constexpr int kN4095 = 4095;
constexpr int kFrame2048 = 2048;
constexpr int kSpecSlots = 65 * 32;  // Hermitian half in (q, r) labels: r = e mod 63 in [0, 32)

// ---------------------------------------------------------------- 4-stage in-place PFA (smem)
PAL_HD constexpr int pfa_inv_mod(int a, int m) {
  for (int t = 1; t < m; ++t)
    if ((a * t) % m == 1) return t;
  return 0;
}
// CRT basis element of factor f: 1 (mod f), 0 (mod 4095/f)
PAL_HD constexpr int pfa_u(int f) { return (kN4095 / f) * pfa_inv_mod((kN4095 / f) % f, f); }
struct Pfa4095 {
  static constexpr int N = kN4095;
  static PAL_HD constexpr int u(int f) { return pfa_u(f); }
  static constexpr int G = (pfa_inv_mod((N / 5) % 5, 5) * pfa_u(5) + pfa_inv_mod((N / 7) % 7, 7) * pfa_u(7) +
                            pfa_inv_mod((N / 9) % 9, 9) * pfa_u(9) + pfa_inv_mod((N / 13) % 13, 13) * pfa_u(13)) % N;
  // after the four in-place stages, output index k sits at location (G*k) mod N
  static PAL_HD constexpr int loc(int k) { return (G * k) % N; }
};
static_assert(pfa_u(5) == 3276 && pfa_u(7) == 1170 && pfa_u(9) == 910 && pfa_u(13) == 2835, "CRT basis");
static_assert(Pfa4095::G == 1829, "output map");

template <int F, int SIGN, typename T, int NT> PAL_DEV void pfa_stage(T* re, T* im) {
  constexpr int U = pfa_u(F);
  for (int g = simt::tid(); g < kN4095 / F; g += NT) {
    T xr[F], xi[F];
    int idx = F * g;
#pragma unroll
    for (int j = 0; j < F; ++j) {
      xr[j] = re[idx];
      xi[j] = im[idx];
      idx += U;
      if (idx >= kN4095) idx -= kN4095;
    }
    dft_odd<F, SIGN, T>(xr, xi);
    idx = F * g;
#pragma unroll
    for (int j = 0; j < F; ++j) {
      re[idx] = xr[j];
      im[idx] = xi[j];
      idx += U;
      if (idx >= kN4095) idx -= kN4095;
    }
  }
}
template <int SIGN, typename T, int NT> PAL_DEV void pfa4095_inplace(T* re, T* im) {
  pfa_stage<13, SIGN, T, NT>(re, im);
  simt::sync_block();
  pfa_stage<9, SIGN, T, NT>(re, im);
  simt::sync_block();
  pfa_stage<7, SIGN, T, NT>(re, im);
  simt::sync_block();
  pfa_stage<5, SIGN, T, NT>(re, im);
  simt::sync_block();
}

// Z = DFT(x1 + i x2) of two real signals -> S1[e], S2[e]  (zL = Z[e], zM = Z[n-e])
template <typename T>
PAL_DEV void split_two_real(T zLr, T zLi, T zMr, T zMi, T& s1r, T& s1i, T& s2r, T& s2i) {
  s1r = T(0.5) * (zLr + zMr);
  s1i = T(0.5) * (zLi - zMi);
  s2r = T(0.5) * (zLi + zMi);
  s2i = T(-0.5) * (zLr - zMr);
}

// PHAT weighting of one cross-spectrum bin, pre-scaled by 1/n (utils.py:116-118)
template <typename T> PAL_DEV void phat_bin(T ar, T ai, T br, T bi, T inv_n, T& rr, T& ri) {
  const T xr = fma_(ar, br, ai * bi);
  const T xi = fma_(ai, br, -(ar * bi));
  const T mag = sqrt_(fma_(xr, xr, xi * xi));
  const T sc = inv_n / (mag + T(1e-10));
  rr = xr * sc;
  ri = xi * sc;
}
// ---------------------------------------------------------------- forward kernel
// grid: one block per (frame, channel pair).  sig: [B][M][2048] float32.
// spec: [B][M][65][32] complex64 -- S[e(r,q)] with e = (2080 r + 2016 q) mod 4095.
// Packed f32x2 arithmetic: the two real channels of a pair travel as (re, im) of one complex
// sequence, one 64-bit shared-memory access and one FFMA2/FADD2 per complex operation.
struct alignas(16) FwdSmem {
  float stage[2][kFrame2048];   // TMA landing zone: the two raw frames
  f2 z[4096];                   // the complex sequence, transformed in place
  mbar_t bar;
  float red[32];                // per-warp partials of the two whitening bounds (NT <= 512)
  float lvl[32];                // per-warp partials of the two channel levels max|x|
  float ssq[32];                // per-warp partials of the two channel energies sum x^2
};

// Exact power-of-two normalisation of a channel to unit level before two channels are packed into one complex
// transform: the rounding residue of the stronger channel (1e-7 of ITS level) would otherwise drown a much weaker
// partner once PHAT whitening brings every bin to unit magnitude.  An all-zero channel gets scale 0 and its spectrum
// is stored as exact zeros, as in the reference (R = 0, corr = 0), instead of the partner's residue.
PAL_DEV void level_scale(float mx, float& sc, float& inv) {
  sc = inv = 0.f;
  if (mx > 0.f && mx < 3.0e38f) {
    int e;
    (void)frexpf(mx, &e);
    e = e < -100 ? -100 : (e > 100 ? 100 : e);
    sc = ldexpf(1.f, -e);
    inv = ldexpf(1.f, e);
  } else if (mx > 0.f) {
    sc = inv = 1.f;
  }
}

// PHAT factorised per CHANNEL instead of per pair.  |Si conj(Sj)| = |Si| |Sj|, so
//     R / (|R| + 1e-10) = Ui conj(Uj) * g,   U = S / |S|,   g = m / (m + 1e-10),  m = |Si| |Sj|      (utils.py:116-117)
// and g is 1 to fp32 precision unless a channel is all but silent.  The forward kernel therefore stores the unit
// phasors U (M normalisations per frame instead of P), the pair kernels multiply two phasors per bin (2 packed
// instructions instead of 8, no MUFU), and the neglected factor is BOUNDED per row:
//     |corr_U[k] - corr[k]| <= (1/n) sum_k 1e-10 / m_k <= 1e-10 sqrt(h_i h_j),   h = (1/n) sum_k 1 / |S_k|^2   (Cauchy-Schwarz)
// h is accumulated here (one float per channel); the pair kernel widens its near-tie margin by the bound and sends
// the row to the float64 kernel when the bound itself is not negligible (frames quieter than about -69 dBFS).
// Bins with S == 0 give R == 0 in both forms and are left out of h.
//
// Rounding noise of the float32 transform.  A bin of the computed spectrum carries noise of about eps32 * rms|S|
// whatever its own size, so the PHASE of a weak bin -- which PHAT then weights like any other -- is off by
// delta_k ~ sigma / |S_k|, sigma = 2^-24 sqrt(sum x^2).  The correlation moves by (1/n) sum_k U_k (e^{i delta_k} - 1) w^{kl}:
// independent errors, standard deviation sqrt(q_i + q_j) / sqrt(n) per sample with q = mean_k sigma^2 / |S_k|^2 = sigma^2 h
// per channel (capped at 4: a phasor cannot move by more than 2) -- no work per bin beyond h.  The pair kernels add
// kNoiseK standard deviations to the near-tie margin and send rows whose VALUES it threatens to the float64 kernel
// (band-limited or tonal frames whose stop band sits below the float32 noise of the pass band).  Calibrated against the
// oracle on the host emulation: rms error / model 0.8 .. 2.3, largest sample error of a row <= 7 model deviations.
constexpr float kNoiseK = 20.f;
constexpr float kEps32Sq = 3.5527137e-15f;      // (2^-24)^2
PAL_DEV f2 whiten_bin(f2 s, float weight, float& hacc) {
  const float x = f2_lo(s), y = f2_hi(s);
  const float m2 = fmaf(x, x, y * y);
  if (m2 > 1e-30f) {
#if PAL_GPU
    float r;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(m2));
#else
    const float r = 1.0f / std::sqrt(m2);
#endif
    hacc = fmaf(weight, r * r, hacc);
    return f2_mul(s, f2_bcast(r));
  }
  if (m2 > 0.f) hacc += weight * 1e30f;     // denormal-range bin: forces the float64 path
  return f2_make(0.f, 0.f);
}

// in-place PFA stage on packed complex data; FIRST: gather the inputs from the raw frames instead
// (z[k] = x0[k] + i x1[k] for k < 2048, zero beyond)
template <int F, int NT, bool FIRST> PAL_DEV void pfa_stage_p(f2* z, const float* x0, const float* x1, bool two, f2 sc01) {
  constexpr int U = pfa_u(F);
  for (int g = simt::tid(); g < kN4095 / F; g += NT) {
    f2 x[F];
    int idx = F * g;
#pragma unroll
    for (int j = 0; j < F; ++j) {
      if (FIRST) x[j] = (idx < kFrame2048) ? f2_mul(f2_make(x0[idx], two ? x1[idx] : 0.f), sc01) : f2_make(0.f, 0.f);
      else x[j] = z[idx];
      idx += U;
      if (idx >= kN4095) idx -= kN4095;
    }
    dft_odd_p<F, -1>(x);
    idx = F * g;
#pragma unroll
    for (int j = 0; j < F; ++j) {
      z[idx] = x[j];
      idx += U;
      if (idx >= kN4095) idx -= kN4095;
    }
  }
}

template <int NT>
PAL_DEV void fwd4095_body(const float* sig, int M, long long n_units /* B * ceil(M/2) */, cpxf* spec, float* hq /* [B][M] */,
                          char* smem_raw) {
  FwdSmem* sm = reinterpret_cast<FwdSmem*>(smem_raw);
  const int tid = simt::tid();
  const int cpairs = (M + 1) / 2;
  if (tid == 0) simt::mbar_init(&sm->bar, 1);
  simt::sync_block();
  unsigned parity = 0;
  // TMA bulk load of the signal frame(s) of one unit: two contiguous 8 KB rows
  auto request = [&](long long u) {
    const long long fr = u / cpairs;
    const int c0 = 2 * int(u % cpairs);
    const bool both = (c0 + 1) < M;
    const float* row0 = sig + (fr * M + c0) * kFrame2048;
    simt::fence_async_smem();
    simt::mbar_expect_tx(&sm->bar, both ? 2u * kFrame2048 * 4u : kFrame2048 * 4u);
    simt::bulk_g2s(sm->stage[0], row0, kFrame2048 * 4u, &sm->bar);
    if (both) simt::bulk_g2s(sm->stage[1], row0 + kFrame2048, kFrame2048 * 4u, &sm->bar);
  };
  if (tid == 0 && simt::bid() < n_units) request(simt::bid());
  for (long long unit = simt::bid(); unit < n_units; unit += simt::nblocks()) {
    const long long frame = unit / cpairs;
    const int ch0 = 2 * int(unit % cpairs);
    const bool two = (ch0 + 1) < M;
    simt::mbar_wait(&sm->bar, parity);
    parity ^= 1u;
    // channel levels -> exact power-of-two scales (level_scale)
    float sc0, sc1, inv0, inv1;
    {
      float m0 = 0.f, m1 = 0.f, e0 = 0.f, e1 = 0.f;
      for (int j = tid; j < kFrame2048 / 4; j += NT) {
        const float4 a = reinterpret_cast<const float4*>(sm->stage[0])[j];
        m0 = fmaxf(m0, fmaxf(fmaxf(fabsf(a.x), fabsf(a.y)), fmaxf(fabsf(a.z), fabsf(a.w))));
        e0 += fmaf(a.x, a.x, a.y * a.y) + fmaf(a.z, a.z, a.w * a.w);
        if (two) {
          const float4 b = reinterpret_cast<const float4*>(sm->stage[1])[j];
          m1 = fmaxf(m1, fmaxf(fmaxf(fabsf(b.x), fabsf(b.y)), fmaxf(fabsf(b.z), fabsf(b.w))));
          e1 += fmaf(b.x, b.x, b.y * b.y) + fmaf(b.z, b.z, b.w * b.w);
        }
      }
#pragma unroll
      for (int m = 16; m >= 1; m >>= 1) {
        m0 = fmaxf(m0, simt::shfl_xor(m0, m));
        m1 = fmaxf(m1, simt::shfl_xor(m1, m));
        e0 += simt::shfl_xor(e0, m);
        e1 += simt::shfl_xor(e1, m);
      }
      if (simt::lane() == 0) {
        sm->lvl[simt::warp()] = m0; sm->lvl[16 + simt::warp()] = m1;
        sm->ssq[simt::warp()] = e0; sm->ssq[16 + simt::warp()] = e1;
      }
      simt::sync_block();
#pragma unroll
      for (int w = 0; w < NT / 32; ++w) { m0 = fmaxf(m0, sm->lvl[w]); m1 = fmaxf(m1, sm->lvl[16 + w]); }
      level_scale(m0, sc0, inv0);
      level_scale(m1, sc1, inv1);
      // (the energy partials stay in shared memory until thread 0 needs them at the end of the unit: nothing more is
      // kept in registers across the four stages, which are tuned to 64 registers)
    }
    pfa_stage_p<13, NT, true>(sm->z, sm->stage[0], sm->stage[1], two, f2_make(sc0, sc1));
    simt::sync_block();
    // the landing zone is free again: the frames of this block's NEXT unit travel while this one is transformed
    if (tid == 0 && unit + simt::nblocks() < n_units) request(unit + simt::nblocks());
    pfa_stage_p<9, NT, false>(sm->z, nullptr, nullptr, two, f2_make(0.f, 0.f));
    simt::sync_block();
    pfa_stage_p<7, NT, false>(sm->z, nullptr, nullptr, two, f2_make(0.f, 0.f));
    simt::sync_block();
    pfa_stage_p<5, NT, false>(sm->z, nullptr, nullptr, two, f2_make(0.f, 0.f));
    simt::sync_block();
    f2* o0 = reinterpret_cast<f2*>(spec + (frame * M + ch0) * kSpecSlots);
    f2* o1 = o0 + kSpecSlots;
    float h0 = 0.f, h1 = 0.f;
    // slot o = (q, r) holds bin e(r, q), which the transform left at location loc(e) = G e mod n; both maps are
    // linear mod n, so the location is (A r + B q) mod n, and a step of NT slots (q += NT / 32, same r) is one
    // add-and-wrap
    static_assert(NT % 32 == 0, "a block step must keep r");
    constexpr int kLocR = (Pfa4095::G * Idx4095::UR) % kN4095, kLocQ = (Pfa4095::G * Idx4095::UQ) % kN4095;
    constexpr int kLocStep = (kLocQ * (NT / 32)) % kN4095;
    int L = (kLocR * (tid & 31) + kLocQ * (tid >> 5)) % kN4095;
    for (int o = tid; o < kSpecSlots; o += NT, L = (L + kLocStep >= kN4095) ? L + kLocStep - kN4095 : L + kLocStep) {
      const int r = o & 31;
      const int L2 = (L == 0) ? 0 : kN4095 - L;
      // Z = DFT(x1 + i x2): S1[e] = (Z[e] + conj(Z[n-e])) / 2, S2[e] = (Z[e] - conj(Z[n-e])) / (2i)
      const f2 zl = sm->z[L], zm = f2_conj(sm->z[L2]);
      const f2 half = f2_bcast(0.5f);
      f2 s0 = f2_mul(f2_add(zl, zm), half);
      const f2 d = f2_mul(f2_sub(zl, zm), half);          // i * S2
      f2 s1 = f2_make(f2_hi(d), -f2_lo(d));               // S2 = -i * d
      if (sc0 == 0.f) s0 = f2_make(0.f, 0.f);             // all-zero channel: exact zeros, not the partner's residue
      if (sc1 == 0.f) s1 = f2_make(0.f, 0.f);
#if PAL_PREWHITEN
      // slot (q, r) stands for bins e and n - e when r > 0; row r = 0 holds both members of a conjugate pair
      const float wgt = r ? 2.f : 1.f;
      s0 = whiten_bin(s0, wgt, h0);
      if (two) s1 = whiten_bin(s1, wgt, h1);
#else
      s0 = f2_mul(s0, f2_bcast(inv0));                    // back to the signal's own level (exact)
      s1 = f2_mul(s1, f2_bcast(inv1));
#endif
      o0[o] = s0;
      if (two) o1[o] = s1;
    }
#if PAL_PREWHITEN
    // both sums in one exchange; the barrier that ends the unit publishes the partials, thread 0 adds them in
    // warp order (deterministic) before it can reach any barrier of the next unit
    h0 = warp_sum(h0);
    h1 = warp_sum(h1);
    if (simt::lane() == 0) { sm->red[simt::warp()] = h0; sm->red[16 + simt::warp()] = h1; }
#endif
    simt::sync_block();
#if PAL_PREWHITEN
    if (tid == 0) {
      float a0 = 0.f, a1 = 0.f, e0 = 0.f, e1 = 0.f;
#pragma unroll
      for (int w = 0; w < NT / 32; ++w) { a0 += sm->red[w]; a1 += sm->red[16 + w]; e0 += sm->ssq[w]; e1 += sm->ssq[16 + w]; }
      a0 *= 1.0f / float(kN4095);
      a1 *= 1.0f / float(kN4095);
      // sigma^2 of the transform's rounding noise at the level the spectra were computed at (see whiten_bin)
      const float sg0 = kEps32Sq * e0 * sc0 * sc0, sg1 = kEps32Sq * e1 * sc1 * sc1;
      // hq[row] = (h, q).  h is wanted at the signal's own level: |S| = |S_scaled| / sc; q = sigma^2 mean(1 / |S_scaled|^2)
      // is level-free (NaN from an overflowed energy ends as 4: the row goes to the float64 kernel)
      float* o = hq + 2 * (frame * M + ch0);
      o[0] = a0 * sc0 * sc0;
      o[1] = (sg0 > 0.f) ? fminf(4.f, sg0 * a0) : 0.f;
      if (two) {
        o[2] = a1 * sc1 * sc1;
        o[3] = (sg1 > 0.f) ? fminf(4.f, sg1 * a1) : 0.f;
      }
    }
#endif
  }
}

// ---------------------------------------------------------------- exact pair kernel (block per pair)
// One complex array, transformed in place twice: channels -> spectra (forward), cross spectrum written over the
// spectra pair by pair (bins e and n-e), cross spectrum -> correlation (inverse).  64 KB in float64, so three
// blocks share an SM.
template <typename T> struct ExactSmem {
  T a_re[4096];
  T a_im[4096];
  unsigned char pkmap[4096];
  PickScratch ps;
  int out_k[16];
  T out_gmax, out_peak;
};

struct PickParams {
  int win_half;   // samples; -1 unbounded; -2 empty window
  int dist;       // int(fs*0.001) >= 1           (utils.py:151)
  int method;     // 0 median, 1 adaptive         (utils.py:144-149)
  float mult;     // threshold_multiplier
  int num_peaks;  // <= 16
};

// items: either all B*P (item_list == nullptr) or the compacted list of flagged items.
// FROM_SPECTRA: read the fp32 spectra written by fwd4095; otherwise transform the two raw
// frames here in precision T (the float64 refinement never sees an fp32-rounded spectrum).
// TS: sample type of `sig` (float: the batched float32 path; double: float64 callers, pal_gcc_phat_tdoa_f64)
template <typename T, int NT, bool FROM_SPECTRA, typename TS = float>
PAL_DEV void pair4095_exact_body(const TS* sig, const cpxf* spec, const int* pairs, int M, int P,
                                 long long n_items, const int* item_list, const int* item_count,
                                 PickParams pp, int* k_idx, int* k_count, float* peak, float* gmax,
                                 unsigned* flags, unsigned extra_flag, unsigned keep_mask, float* corr_out,
                                 char* smem_raw) {
  ExactSmem<T>* sm = reinterpret_cast<ExactSmem<T>*>(smem_raw);
  const int tid = simt::tid();
  const long long total = item_list ? (long long)(*item_count) : n_items;
  const T inv_n = T(1) / T(kN4095);
  for (long long it = simt::bid(); it < total; it += simt::nblocks()) {
    const long long item = item_list ? (long long)item_list[it] : it;
    const long long frame = item / P;
    const int p = int(item % P);
    const int mi = pairs[2 * p], mj = pairs[2 * p + 1];
    // The in-place prime-factor transform leaves bin e at location loc(e) = G e mod n.  Feeding an array in THAT
    // order to the transform again returns its result in natural order (DFT of x[G^-1 m] is X[G k], and G k is
    // exactly where the transform puts output k), so the cross spectrum is written where the spectra already are
    // and the correlation row comes out at a_re[k] -- no permutation pass, no second array.
    if (FROM_SPECTRA) {
      const cpxf* si = spec + (frame * M + mi) * kSpecSlots;
      const cpxf* sj = spec + (frame * M + mj) * kSpecSlots;
      for (int o = tid; o < kSpecSlots; o += NT) {
        const int q = o >> 5, r = o & 31;
        const int e = Idx4095::elem(r, q);
        const cpxf a = si[o], b = sj[o];
        T rr, ri;
        phat_bin<T>(T(a.x), T(a.y), T(b.x), T(b.y), inv_n, rr, ri);
        const int L = Pfa4095::loc(e);
        sm->a_re[L] = rr;
        sm->a_im[L] = ri;
        if (e != 0) { sm->a_re[kN4095 - L] = rr; sm->a_im[kN4095 - L] = -ri; }      // loc(n - e) = n - loc(e)
      }
    } else {
      const TS* xi = sig + (frame * M + mi) * kFrame2048;
      const TS* xj = sig + (frame * M + mj) * kFrame2048;
      for (int k = tid; k < 4096; k += NT) {
        sm->a_re[k] = (k < kFrame2048) ? T(xi[k]) : T(0);
        sm->a_im[k] = (k < kFrame2048) ? T(xj[k]) : T(0);
      }
      simt::sync_block();
      pfa4095_inplace<-1, T, NT>(sm->a_re, sm->a_im);
      // R[e] and R[n-e] = conj(R[e]) both come from Z[e] and Z[n-e] (at L and n - L): each thread owns such
      // pairs and overwrites them
      for (int e = tid; e <= kN4095 / 2; e += NT) {
        const int L = Pfa4095::loc(e);
        const int L2 = (L == 0) ? 0 : kN4095 - L;
        T s1r, s1i, s2r, s2i;
        split_two_real<T>(sm->a_re[L], sm->a_im[L], sm->a_re[L2], sm->a_im[L2], s1r, s1i, s2r, s2i);
        T rr, ri;
        phat_bin<T>(s1r, s1i, s2r, s2i, inv_n, rr, ri);
        sm->a_re[L] = rr;
        sm->a_im[L] = ri;
        if (L2 != L) { sm->a_re[L2] = rr; sm->a_im[L2] = -ri; }
      }
    }
    simt::sync_block();
    pfa4095_inplace<+1, T, NT>(sm->a_re, sm->a_im);
    T* const crow = sm->a_re;
    if (corr_out)
      for (int k = tid; k < kN4095; k += NT) corr_out[item * kN4095 + k] = float(crow[k]);
    PickResult pr = peakpick_row<T, NT>(crow, kN4095, kFrame2048 - 1, pp.win_half, pp.dist, pp.method,
                                        T(pp.mult), pp.num_peaks, sm->pkmap, &sm->ps, sm->out_k,
                                        &sm->out_gmax, &sm->out_peak);
    simt::sync_block();
    if (tid == 0) {
      for (int t = 0; t < pp.num_peaks; ++t) k_idx[item * pp.num_peaks + t] = (t < pr.count) ? sm->out_k[t] : -1;
      if (k_count) k_count[item] = pr.count;
      peak[item] = float(sm->out_peak);
      gmax[item] = float(sm->out_gmax);
      // keep_mask != 0: this is a re-evaluation; remember why the fast path asked for it
      flags[item] = (keep_mask ? (flags[item] & keep_mask) : 0u) | pr.flags | extra_flag;
    }
    simt::sync_block();
  }
}

// ---------------------------------------------------------------- fast pair kernel (warp per pair)
// All complex arithmetic is packed (re, im) f32x2 (FFMA2 / FADD2 / FMUL2, pal_simt.h): half the
// issue slots of scalar code.  The correlation row is kept UNSCALED (n * corr): the 1/n of the
// inverse DFT is applied only to the two values that leave the kernel.
//
// Per-warp shared memory: the phase-A -> phase-B exchange Y[r][kq] (r < 32 by Hermitian symmetry),
// later overwritten by the 4095-sample correlation row the peak pick scans.
struct alignas(16) FastWarpSmem {
  union {
    struct {
      f2 ya[32 * 33];   // Y[r][kq], kq = 1..32  -> column kq-1
      f2 yb[32 * 33];   // Y[r][kq], kq = 33..64 -> column kq-33
      f2 y0[32];        // Y[r][0]
    } y;
    float corr[4096];
  };
  f2 lx[64];            // exchange of the odd column kq = 0 (DFT-7 -> DFT-9)
};

constexpr float kNegBig = -3.0e38f;

// PHAT weighting of one cross-spectrum bin a * conj(b) / (|a * conj(b)| + 1e-10)  (utils.py:116-117),
// 8 issue slots: FMUL2, FFMA2 (swizzled operand), FMUL, FFMA, MUFU.SQRT, FADD, MUFU.RCP, FMUL2.
// The approximate sqrt / reciprocal (about 1 ulp each) touch only the MAGNITUDE of the weight
// (2e-7 relative), never the phase, so the correlation moves far less than the 1e-4 tolerance.
PAL_DEV f2 phat_bin_p(f2 a, f2 b) {
  const f2 t = f2_mul(a, f2_bcast(f2_lo(b)));                                    // (ar br, ai br)
  const f2 x = f2_fma(f2_make(f2_hi(a), -f2_lo(a)), f2_bcast(f2_hi(b)), t);      // + (ai bi, -ar bi)
#if PAL_PREWHITEN
  return x;      // a, b are unit phasors (whiten_bin): the product is the PHAT-weighted bin
#else
  const float xr = f2_lo(x), xi = f2_hi(x);
  const float m2 = fmaf(xr, xr, xi * xi);
#if PAL_GPU
  float mag, rcp;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(mag) : "f"(m2));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rcp) : "f"(mag + 1e-10f));
#else
  const float rcp = 1.0f / (std::sqrt(m2) + 1e-10f);
#endif
  return f2_mul(x, f2_bcast(rcp));
#endif
}

// Bound (in the unscaled row, n * corr) on what the fast path neglects for pair (i, j): the factor g of the whitened
// spectra, n * 1e-10 * sqrt(h_i h_j), plus kNoiseK standard deviations of the transform's rounding noise,
// n * kNoiseK * sqrt((q_i + q_j) / n); see whiten_bin.  hq: [rows][2] = (h, q).
PAL_DEV float whiten_bound(const float* hq, long long row_i, long long row_j) {
#if PAL_PREWHITEN
  const cpxf a = *reinterpret_cast<const cpxf*>(hq + 2 * row_i), b = *reinterpret_cast<const cpxf*>(hq + 2 * row_j);
  return (1e-10f * float(kN4095)) * sqrt_(a.x * b.x) + kNoiseK * sqrt_(float(kN4095) * (a.y + b.y));
#else
  return 0.f;
#endif
}

// Careful scan of the window (strict local maxima, runner-up, equal neighbours).  Only used when
// the plain maximum of the window is not itself a strict local maximum (window edge / plateau).
PAL_DEV void window_scan_slow(const float* c, int lo, int hi, int lane, float& bv, int& bi, float& cand2,
                              float& pl) {
  float b1 = 0.f, b2 = kNegBig;
  pl = kNegBig;
  int i1 = -1;
  for (int k = lo + lane; k <= hi; k += 32) {
    const float v = c[k], vl = c[k - 1], vr = c[k + 1];
    if (v == vl || v == vr) pl = fmaxf(pl, v);
    if (vl < v && v > vr) {
      if (i1 < 0 || v >= b1) { if (i1 >= 0) b2 = fmaxf(b2, b1); b1 = v; i1 = k; }
      else b2 = fmaxf(b2, v);
    }
  }
  bv = b1;
  bi = i1;
  warp_argmax<true>(bv, bi);
  cand2 = (i1 == bi) ? b2 : ((i1 >= 0) ? b1 : kNegBig);
#pragma unroll
  for (int m = 16; m >= 1; m >>= 1) {
    cand2 = fmaxf(cand2, simt::shfl_xor(cand2, m));
    pl = fmaxf(pl, simt::shfl_xor(pl, m));
  }
}

// warp-wide max in ONE instruction (Blackwell CREDUX.MAX.F32 / .S32, redux.sync, sm_100a)
PAL_DEV float warp_max_f32(float v) {
#if PAL_GPU
  float r;
  asm volatile("redux.sync.max.f32 %0, %1, 0xffffffff;" : "=f"(r) : "f"(v));
  return r;
#else
#pragma unroll
  for (int m = 16; m >= 1; m >>= 1) v = fmaxf(v, simt::shfl_xor(v, m));
  return v;
#endif
}
PAL_DEV int warp_max_s32(int v) {
#if PAL_GPU
  int r;
  asm volatile("redux.sync.max.s32 %0, %1, 0xffffffff;" : "=r"(r) : "r"(v));
  return r;
#else
#pragma unroll
  for (int m = 16; m >= 1; m >>= 1) { const int o = simt::shfl_xor(v, m); v = o > v ? o : v; }
  return v;
#endif
}

// Everything the peak pick of the fast kernels needs besides the row itself.
struct FastPick {
  int lo, hi, g_lo, g_hi, dist;
  float eps_s, mean_bound, inv_n;
};
PAL_DEV FastPick make_fast_pick(int win_half, int dist, float eps) {
  FastPick k;
  k.inv_n = 1.0f / float(kN4095);
  k.eps_s = eps * float(kN4095);       // tolerance in the unscaled row
  // mean|c| <= rms(c) <= 1/sqrt(n) by Parseval (|R| <= 1 after PHAT); in the unscaled row: sqrt(n).
  // A winner above this bound is above mean|c| whatever the row looks like (utils.py:155,166).
  k.mean_bound = 64.0f;                // sqrt(4095) = 63.99 rounded up
  const int c0 = kFrame2048 - 1;
  k.lo = 1; k.hi = kN4095 - 2;
  if (win_half >= 0) {
    k.lo = (c0 - win_half > 1) ? c0 - win_half : 1;
    k.hi = (c0 + win_half < kN4095 - 2) ? c0 + win_half : kN4095 - 2;
  } else if (win_half < -1) {
    k.lo = 1; k.hi = 0;
  }
  // window split into 16-byte groups [g_lo, g_hi) plus at most 3 + 3 edge samples
  k.g_lo = (k.lo + 3) & ~3;
  k.g_hi = (k.hi + 1) & ~3;
  k.dist = dist;
  return k;
}

// Peak pick (num_peaks = 1) of one unscaled correlation row in shared memory by one warp.
template <bool WRITE_CORR>
PAL_DEV void fast_pick_row(const float* c, long long item, float gm, int lane, const FastPick& pk, float wbound,
                           int* k_idx, float* peak, float* gmax, unsigned* flags, float* corr_out) {
  const int lo = pk.lo, hi = pk.hi, g_lo = pk.g_lo, g_hi = pk.g_hi, dist = pk.dist;
  gm = warp_max_f32(gm);
  // wbound: see whiten_bound.  Small: it widens the near-tie margin.  Not small against the tie margin and against
  // the row's maximum (or NaN/inf): the row's VALUES may be off by a third of the 1e-4 tolerance (the bound is about 3 x the largest error seen) ->
  // float64 kernel.
  const bool quiet = !(wbound <= 2.f * pk.eps_s + 1e-4f * fmaxf(gm, 0.f));
  const float eps_s = pk.eps_s + (quiet ? 0.f : wbound), mean_bound = pk.mean_bound, inv_n = pk.inv_n;
  if (WRITE_CORR) {
    float* dst = corr_out + item * kN4095;
    for (int k = lane; k < kN4095; k += 32) dst[k] = c[k] * inv_n;
  }

  // largest and second largest SAMPLE of the window (vectorised, no neighbour tests): the
  // largest one is the answer whenever it is a strict local maximum, which is then verified
  float b1 = kNegBig, b2 = kNegBig;
  int gsel = -1;
  for (int g = g_lo + 4 * lane; g < g_hi; g += 128) {
    const float4 v = *reinterpret_cast<const float4*>(c + g);
    const float m01 = fmaxf(v.x, v.y), n01 = fminf(v.x, v.y);
    const float m23 = fmaxf(v.z, v.w), n23 = fminf(v.z, v.w);
    const float m = fmaxf(m01, m23);
    const float s4 = fmaxf(fminf(m01, m23), fmaxf(n01, n23));     // second largest of the four
    b2 = fmaxf(b2, fmaxf(fminf(b1, m), s4));
    if (m > b1) gsel = g;
    b1 = fmaxf(b1, m);
  }
  int ksel = -1;
  if (gsel >= 0) {
    const float4 v = *reinterpret_cast<const float4*>(c + gsel);
    ksel = gsel + ((v.w == b1) ? 3 : (v.z == b1) ? 2 : (v.y == b1) ? 1 : 0);
  }
  if (lane < 6) {                      // the <= 3 + 3 samples outside the aligned groups
    const int k = (lane < 3) ? lo + lane : g_hi + (lane - 3);
    const bool ok = (lane < 3) ? (k < g_lo && k <= hi) : (k <= hi && g_hi >= g_lo);
    if (ok) {
      const float v = c[k];
      b2 = fmaxf(b2, fminf(b1, v));
      if (v > b1) ksel = k;
      b1 = fmaxf(b1, v);
    }
  }
  float bv = warp_max_f32(b1);
  int bi = warp_max_s32((b1 == bv) ? ksel : -1);
  float cand2 = warp_max_f32((ksel == bi) ? b2 : b1);
  float pl = kNegBig;
  if (bi >= 0 && !(c[bi - 1] < bv && bv > c[bi + 1])) window_scan_slow(c, lo, hi, lane, bv, bi, cand2, pl);

  unsigned fl = quiet ? PAL_FLAG_NEAR_TIE : 0u;
  int kbest;
  float hbest;
  if (bi >= 0 && bv >= mean_bound + eps_s) {
    kbest = bi;
    hbest = bv;
    if (cand2 >= bv - eps_s) fl |= PAL_FLAG_NEAR_TIE;
    if (pl >= bv - eps_s) fl |= PAL_FLAG_PLATEAU;
    // Anything within `dist` samples that is as high as the winner (up to eps) needs the full
    // greedy resolution of the distance rule -> exact kernel.  Samples inside the window are
    // already covered by cand2; only a winner close to a window edge has neighbours outside.
    if (bi - dist < lo || bi + dist > hi) {
      bool hit = false;
      for (int o = -dist + lane; o <= dist; o += 32) {
        const int q = bi + o;
        if (o != 0 && q >= 0 && q < kN4095 && c[q] >= bv - eps_s) hit = true;
      }
      if (simt::ballot(hit)) fl |= PAL_FLAG_CHAIN;
    }
  } else if (bi >= 0) {
    // winner not clearly above mean|c|: the median / fallback logic decides -> exact kernel
    kbest = bi;
    hbest = bv;
    fl |= PAL_FLAG_NEAR_TIE;
  } else {
    // no local maximum in the window: unbounded first argmax (utils.py:168-172)
    int gi = 0x7fffffff;
    int near = 0;
    bool nonzero = false;
    for (int k = lane; k < kN4095; k += 32) {
      const float v = c[k];
      if (v == gm && k < gi) gi = k;
      if (v >= gm - eps_s) ++near;
      nonzero |= (v != 0.f);
    }
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) {
      const int og = simt::shfl_xor(gi, m);
      gi = og < gi ? og : gi;
      near += simt::shfl_xor(near, m);
    }
    kbest = gi;
    hbest = gm;
    fl |= PAL_FLAG_FALLBACK_ARGMAX;
    const bool all_zero = simt::ballot(nonzero) == 0u;
    if (near > 1 && !all_zero) fl |= PAL_FLAG_NEAR_TIE;
    if (pl > kNegBig && !all_zero) fl |= PAL_FLAG_PLATEAU;
  }
  if (lane == 0) {
    k_idx[item] = kbest;
    peak[item] = hbest * inv_n;
    gmax[item] = gm * inv_n;
    flags[item] = fl;
  }
}

// Bins of the NEXT pair whose spectra are fetched into registers before the peak pick of the current
// pair starts (the pick needs few registers and ~2000 cycles: it hides the L2 latency completely).
#ifndef PAL_PREFETCH_BINS
#define PAL_PREFETCH_BINS 8
#endif
#ifndef PAL_L1_PREFETCH_BINS
#define PAL_L1_PREFETCH_BINS 8    // = kPrefetchBins: off (measured: no gain on B200, profiles/)
#endif
constexpr int kPrefetchBins = PAL_PREFETCH_BINS;        // multiple of 4
constexpr int kL1PrefetchBins = PAL_L1_PREFETCH_BINS;   // bins [kPrefetchBins, kL1PrefetchBins) go to L1 only

template <int WARPS, bool WRITE_CORR>
PAL_DEV void pair4095_fast_body(const cpxf* spec, const float* hq, const int* pairs, int M, int P, long long n_items,
                                int win_half, int dist, float eps, int* k_idx, float* peak, float* gmax,
                                unsigned* flags, float* corr_out, char* smem_raw) {
  using P65 = Pfa2<5, 13>;
  using P63 = Pfa2<7, 9>;
  const int lane = simt::lane();
#ifdef PAL_UNIFORM_WARP
  const int warp = simt::shfl(simt::warp(), 0);     // provably warp-uniform (see pair4095_tmem_body)
#else
  const int warp = simt::warp();
#endif
  FastWarpSmem* sm = reinterpret_cast<FastWarpSmem*>(smem_raw) + warp;
  const FastPick pk = make_fast_pick(win_half, dist, eps);
  // scatter bases: output k = (63 kq + 65 kr) mod 4095 of columns kq = lane+1 and lane+33;
  // p?w are the same bases pre-wrapped by -4095 so that every store is [base + immediate]
  const int b0 = 63 * (lane + 1);
  float* const p1 = sm->corr + b0;
  float* const p1w = p1 - kN4095;
  float* const p2 = p1 + 63 * 32;
  float* const p2w = p2 - kN4095;

  const long long stride = (long long)simt::nblocks() * WARPS;
  long long item = (long long)simt::bid() * WARPS + warp;
  if (item >= n_items) return;
#if PAL_GPU && PAL_STAGGER > 0
  // De-phase the two warps that share an SM sub-partition (warp w and w + WARPS/2) once, at kernel
  // start.  Measured on B200: 7 % faster at 16384 frames (bursts to L2 no longer line up); the
  // amount does not matter (profiles/).
  if (warp >= WARPS / 2) {
    const long long t0 = clock64();
    while (clock64() - t0 < (long long)(PAL_STAGGER)) {}
  }
#endif

  // spectrum rows of the current pair, and the first kPrefetchBins bins of both in registers
  const f2* si;
  const f2* sj;
  f2 pa[kPrefetchBins], pb[kPrefetchBins];
  float wb, wb_next = 0.f;     // whiten_bound of the current / next pair
  long long frame_it = item / P;          // (frame, pair) of the item in flight: add-and-carry per step, no division
  int pair_it = int(item - frame_it * P);
  const long long stride_f = stride / P;
  const int stride_p = int(stride - stride_f * P);
  {
    const long long frame = frame_it;
    const int p = pair_it;
    si = reinterpret_cast<const f2*>(spec + (frame * M + pairs[2 * p]) * kSpecSlots) + lane;
    sj = reinterpret_cast<const f2*>(spec + (frame * M + pairs[2 * p + 1]) * kSpecSlots) + lane;
    wb = whiten_bound(hq, frame * M + pairs[2 * p], frame * M + pairs[2 * p + 1]);
#pragma unroll
    for (int q = 0; q < kPrefetchBins; ++q) { pa[q] = si[q * 32]; pb[q] = sj[q * 32]; }
  }

  for (;;) {
    float gm = kNegBig;
    {
      // ---- phase A: lane r owns elements e(r, q), q = 0..64: PHAT, then DFT-65 over q as a 5 x 13
      // prime-factor transform.  Pass 1 consumes the bins five at a time (PHAT + DFT-5), pass 2
      // stores each DFT-13 group as soon as it is done, so loads, MUFU work and shared-memory
      // stores are spread over the arithmetic instead of arriving in bursts.
      f2 z[65];
#pragma unroll
      for (int b = 0; b < 13; ++b) {
        f2 t[5];
#pragma unroll
        for (int a = 0; a < 5; ++a) {
          const int q = P65::slot(a, b);
          t[a] = (q < kPrefetchBins) ? phat_bin_p(pa[q < kPrefetchBins ? q : 0], pb[q < kPrefetchBins ? q : 0])
                                     : phat_bin_p(si[q * 32], sj[q * 32]);
        }
        dft_odd_p<5, +1>(t);
#pragma unroll
        for (int a = 0; a < 5; ++a) z[P65::slot(a, b)] = t[a];
      }
#pragma unroll
      for (int a = 0; a < 5; ++a) {
        f2 t[13];
#pragma unroll
        for (int b = 0; b < 13; ++b) t[b] = z[P65::slot(a, b)];
        dft_odd_p<13, +1>(t);
#pragma unroll
        for (int b = 0; b < 13; ++b) {
          const int kq = P65::out_index(P65::slot(a, b));
          if (kq == 0) sm->y.y0[lane] = t[b];
          else if (kq <= 32) sm->y.ya[lane * 33 + (kq - 1)] = t[b];
          else sm->y.yb[lane * 33 + (kq - 33)] = t[b];
        }
      }
    }
    simt::sync_warp();
    {
      // ---- phase B: lane l owns output columns kq = l+1 (real part) and l+33 (imaginary part)
      // of ONE complex DFT-63: w[r] = Y[r][l+1] + i Y[r][l+33]; every column is Hermitian in r
      // (Y[63-r][kq] = conj(Y[r][kq])), so only r < 32 was computed in phase A
      f2 w[63];
#pragma unroll
      for (int r = 0; r < 32; ++r) {
        const f2 a = sm->y.ya[r * 33 + lane], b = sm->y.yb[r * 33 + lane];
        if (r == 0) {
          w[0] = f2_make(f2_lo(a), f2_lo(b));
        } else {
          w[r] = f2_add(a, f2_muli(b));                   // (a.x - b.y, a.y + b.x)
          w[63 - r] = f2_add(f2_conj(a), f2_swap(b));     // (a.x + b.y, b.x - a.y)
        }
      }
      // odd column kq = 0: nine DFT-7 on lanes 0..8 now, seven DFT-9 on lanes 0..6 below
      f2 t7[7];
      if (lane < 9) {
#pragma unroll
        for (int a = 0; a < 7; ++a) {
          const int r = (a * P63::UA + lane * P63::UB) % 63;
          const f2 v = sm->y.y0[r < 32 ? r : 63 - r];
          t7[a] = f2_make(f2_lo(v), (r == 0) ? 0.f : (r < 32 ? f2_hi(v) : -f2_hi(v)));
        }
      }
      simt::sync_warp();            // every read of Y is done: the union now holds the correlation row
      if (lane < 9) {
        dft_odd_p<7, +1>(t7);
#pragma unroll
        for (int a = 0; a < 7; ++a) sm->lx[(a * P63::UA + lane * P63::UB) % 63] = t7[a];
      }
      // DFT-63 as a 7 x 9 prime-factor transform; each DFT-9 group is scattered to natural order as
      // soon as it is done: consecutive lanes are 63 floats apart -> conflict-free.  Whether
      // (63 kq + 65 kr) wraps past 4095 is known at compile time for half of the stores; max(c) is
      // taken from the registers on the way (FMNMX3).
#pragma unroll
      for (int b = 0; b < 9; ++b) {
        f2 t[7];
#pragma unroll
        for (int a = 0; a < 7; ++a) t[a] = w[P63::slot(a, b)];
        dft_odd_p<7, +1>(t);
#pragma unroll
        for (int a = 0; a < 7; ++a) w[P63::slot(a, b)] = t[a];
      }
#pragma unroll
      for (int a = 0; a < 7; ++a) {
        f2 t[9];
#pragma unroll
        for (int b = 0; b < 9; ++b) t[b] = w[P63::slot(a, b)];
        dft_odd_p<9, +1>(t);
#pragma unroll
        for (int b = 0; b < 9; ++b) {
          const int kr = P63::out_index(P63::slot(a, b));
          const int c = 65 * kr;
          const float vr = f2_lo(t[b]), vi = f2_hi(t[b]);
          if (kr <= 31) p1[c] = vr;
          else ((b0 + c >= kN4095) ? p1w : p1)[c] = vr;
          if (kr == 0) p2[c] = vi;
          else if (kr >= 32) p2w[c] = vi;
          else ((b0 + 63 * 32 + c >= kN4095) ? p2w : p2)[c] = vi;
          gm = fmaxf(gm, fmaxf(vr, vi));
        }
      }
      simt::sync_warp();
      if (lane < 7) {
        f2 u[9];
#pragma unroll
        for (int b = 0; b < 9; ++b) u[b] = sm->lx[(lane * P63::UA + b * P63::UB) % 63];
        dft_odd_p<9, +1>(u);
#pragma unroll
        for (int b = 0; b < 9; ++b) {
          const int kr = (9 * lane + 7 * b) % 63;
          const float v = f2_lo(u[b]);
          sm->corr[(65 * kr) % kN4095] = v;
          gm = fmaxf(gm, v);
        }
      }
    }
    // ---- fetch the first bins of the NEXT pair while this one is being picked -----------------
    const long long next = item + stride;
    const bool has_next = next < n_items;
    if (has_next) {
      pair_it += stride_p;
      frame_it += stride_f;
      if (pair_it >= P) { pair_it -= P; ++frame_it; }
      const long long frame = frame_it;
      const int p = pair_it;
      si = reinterpret_cast<const f2*>(spec + (frame * M + pairs[2 * p]) * kSpecSlots) + lane;
      sj = reinterpret_cast<const f2*>(spec + (frame * M + pairs[2 * p + 1]) * kSpecSlots) + lane;
      wb_next = whiten_bound(hq, frame * M + pairs[2 * p], frame * M + pairs[2 * p + 1]);
#pragma unroll
      for (int q = 0; q < kPrefetchBins; ++q) { pa[q] = si[q * 32]; pb[q] = sj[q * 32]; }
#if PAL_GPU
      // ... and pull the following bins into L1 (CCTL.PF1, no registers): 32 lanes x 32-byte sectors
      // per instruction = 4 bins of one row
#pragma unroll
      for (int t = kPrefetchBins / 4; t < kL1PrefetchBins / 4; ++t) {
        asm volatile("prefetch.global.L1 [%0];" ::"l"(reinterpret_cast<const char*>(si - lane) + (t * 32 + lane) * 32));
        asm volatile("prefetch.global.L1 [%0];" ::"l"(reinterpret_cast<const char*>(sj - lane) + (t * 32 + lane) * 32));
      }
#endif
    }
    simt::sync_warp();

    // ---- peak pick (num_peaks = 1), utils.py:140-181 in the reduced form ----------------
    fast_pick_row<WRITE_CORR>(sm->corr, item, gm, lane, pk, wb, k_idx, peak, gmax, flags, corr_out);
    if (!has_next) break;
    item = next;
    wb = wb_next;
    simt::sync_warp();   // the next item overwrites the union
  }
}

// ---------------------------------------------------------------- fast pair kernel, TMEM-assisted
// Same algorithm as pair4095_fast_body, but the register tiles between the two passes of each
// prime-factor transform (65 and 63 complex values per lane) are parked in Tensor Memory
// (lane-private columns, tcgen05.st / tcgen05.ld) instead of staying in registers.  The working
// set drops from 255 to <= 168 registers per thread, so 12 warps (3 per scheduler) are resident
// per SM instead of 8, which is what hides the latency of the exchange / scatter / pick phases.
// TMEM columns of a warp: [colbase, colbase + 130).
template <int WARPS> struct TmemPlan {
  static constexpr int kSlots = (WARPS + 3) / 4;                 // warps per lane quadrant
  static constexpr int kColsPerWarp = ((512 / kSlots) / 2) * 2;
  static_assert(kColsPerWarp >= 130, "not enough tensor-memory columns per warp");
};

template <int WARPS, bool WRITE_CORR>
PAL_DEV void pair4095_tmem_body(const cpxf* spec, const float* hq, const int* pairs, int M, int P, long long n_items,
                                int win_half, int dist, float eps, int* k_idx, float* peak, float* gmax,
                                unsigned* flags, float* corr_out, char* smem_raw, int pairs_in_smem = 0) {
  using P65 = Pfa2<5, 13>;
  using P63 = Pfa2<7, 9>;
  const int lane = simt::lane();
  // The warp index is broadcast from lane 0 so that the compiler can prove it warp-uniform: every
  // branch on the work item is then a uniform branch, the main loop is convergent code and
  // descriptors / tensor-memory addresses stay in uniform registers.
  const int warp = simt::shfl(simt::warp(), 0);
  FastWarpSmem* sm = reinterpret_cast<FastWarpSmem*>(smem_raw) + warp;
  unsigned* tmem_slot = reinterpret_cast<unsigned*>(reinterpret_cast<FastWarpSmem*>(smem_raw) + WARPS);
  // The microphone indices of the NEXT pair steer thirty spectrum loads; read from global memory they cost a full
  // L2 round trip in front of those loads (2 % of the kernel's stall samples).  When the table fits behind the warp
  // tiles (pairs_in_smem: P <= ~2500) it is copied there once per block.
  int* const pairs_s = reinterpret_cast<int*>(tmem_slot + 4);
  if (pairs_in_smem)
    for (int i = simt::tid(); i < 2 * P; i += WARPS * 32) pairs_s[i] = pairs[i];
  if (warp == 0) simt::tmem_alloc512(tmem_slot);
  simt::tmem_fence_before_sync();
  simt::sync_block();
  simt::tmem_fence_after_sync();
  const unsigned tbase = *tmem_slot;
  const unsigned tcol = simt::tmem_addr(tbase, warp, (warp >> 2) * TmemPlan<WARPS>::kColsPerWarp);

  const FastPick pk = make_fast_pick(win_half, dist, eps);
  const int b0 = 63 * (lane + 1);
  float* const p1 = sm->corr + b0;
  float* const p1w = p1 - kN4095;
  float* const p2 = p1 + 63 * 32;
  float* const p2w = p2 - kN4095;

  const long long stride = (long long)simt::nblocks() * WARPS;
  long long item = (long long)simt::bid() * WARPS + warp;
  // The first pair's microphone indices are fetched here, in convergent code, on purpose: it makes
  // the compiler set up the global-memory descriptor in a uniform register once; if the first global
  // load sits inside the (formally divergent) block below, every one of the 130 spectrum loads of a
  // pair pays two extra R2UR instructions.
  const long long item_c = item < n_items ? item : n_items - 1;
  const int mi0 = pairs[2 * int(item_c % P)], mj0 = pairs[2 * int(item_c % P) + 1];
  if (item < n_items) {
#if PAL_GPU && PAL_STAGGER > 0
    if (warp >= 4) {   // de-phase the warps that share a scheduler
      const long long t0 = clock64();
      while (clock64() - t0 < (long long)(PAL_STAGGER) * (warp >> 2)) {}
    }
#endif
    // Software pipeline of the spectrum loads: the bins of DFT-5 group b + kLA are requested while
    // group b is processed (a ring of kLA groups of 5 + 5 complex values in registers); the first
    // kLA groups of the NEXT pair are requested before the peak pick of the current one.
#ifndef PAL_TMEM_LOOKAHEAD
#define PAL_TMEM_LOOKAHEAD 3
#endif
    constexpr int kLA = PAL_TMEM_LOOKAHEAD;
    const f2* si;
    const f2* sj;
    f2 ra[kLA][5], rb[kLA][5];
    float wb, wb_next = 0.f;     // whiten_bound of the current / next pair
    // (frame, pair) of the item in flight, advanced by the constant stride with an add-and-carry instead of a
    // 64-bit division per pair
    long long frame_it = item / P;
    int pair_it = int(item - frame_it * P);
    const long long stride_f = stride / P;
    const int stride_p = int(stride - stride_f * P);
    {
      const long long frame = frame_it;
      si = reinterpret_cast<const f2*>(spec + (frame * M + mi0) * kSpecSlots) + lane;
      sj = reinterpret_cast<const f2*>(spec + (frame * M + mj0) * kSpecSlots) + lane;
      wb = whiten_bound(hq, frame * M + mi0, frame * M + mj0);
#pragma unroll
      for (int g = 0; g < kLA; ++g)
#pragma unroll
        for (int a5 = 0; a5 < 5; ++a5) { ra[g][a5] = si[P65::slot(a5, g) * 32]; rb[g][a5] = sj[P65::slot(a5, g) * 32]; }
    }
    for (;;) {
      float gm = kNegBig;
      // ---- phase A, pass 1: PHAT + DFT-5 over a for every b; result (a, b) parked at column 2 (13 a + b)
#pragma unroll
      for (int b = 0; b < 13; ++b) {
        f2 t[5];
#pragma unroll
        for (int a = 0; a < 5; ++a) t[a] = phat_bin_p(ra[b % kLA][a], rb[b % kLA][a]);
        if (b + kLA < 13) {
#pragma unroll
          for (int a = 0; a < 5; ++a) {
            ra[b % kLA][a] = si[P65::slot(a, b + kLA) * 32];
            rb[b % kLA][a] = sj[P65::slot(a, b + kLA) * 32];
          }
        }
        dft_odd_p<5, +1>(t);
#pragma unroll
        for (int a = 0; a < 5; ++a) {
          const f2 one[1] = {t[a]};
          tmem_st(tcol + 2 * (13 * a + b), one);
        }
      }
      simt::tmem_wait_st();
      // ---- phase A, pass 2: DFT-13 over b for every a, straight into the exchange tile
#pragma unroll
      for (int a = 0; a < 5; ++a) {
        f2 t[13];
        {
          f2 v8[8], v4[4], v1[1];
          tmem_ld(tcol + 26 * a, v8);
          tmem_ld(tcol + 26 * a + 16, v4);
          tmem_ld(tcol + 26 * a + 24, v1);
          simt::tmem_wait_ld();
#pragma unroll
          for (int i = 0; i < 8; ++i) t[i] = v8[i];
#pragma unroll
          for (int i = 0; i < 4; ++i) t[8 + i] = v4[i];
          t[12] = v1[0];
        }
        dft_odd_p<13, +1>(t);
#pragma unroll
        for (int b = 0; b < 13; ++b) {
          const int kq = P65::out_index(P65::slot(a, b));
          if (kq == 0) sm->y.y0[lane] = t[b];
          else if (kq <= 32) sm->y.ya[lane * 33 + (kq - 1)] = t[b];
          else sm->y.yb[lane * 33 + (kq - 33)] = t[b];
        }
      }
      simt::sync_warp();
      // ---- phase B, pass 1: the DFT-7 groups b and 9-b together (they share the rows r and 63-r of
      // the Hermitian columns); result (a, b) parked at column 2 (9 a + b)
#pragma unroll
      for (int b = 0; b < 5; ++b) {
        const int bc = (9 - b) % 9;
        f2 t1[7], t2[7];
#pragma unroll
        for (int a = 0; a < 7; ++a) {
          const int s = P63::slot(a, b);
          const int ac = (7 - a) % 7;                       // slot(ac, bc) == (63 - s) % 63
          if (s == 0) {
            const f2 ya = sm->y.ya[lane], yb = sm->y.yb[lane];
            t1[0] = f2_make(f2_lo(ya), f2_lo(yb));
          } else if (b != 0 || s < 32) {
            const int r = s < 32 ? s : 63 - s;
            const f2 ya = sm->y.ya[r * 33 + lane], yb = sm->y.yb[r * 33 + lane];
            const f2 wlo = f2_add(ya, f2_muli(yb));                 // w[r]
            const f2 whi = f2_add(f2_conj(ya), f2_swap(yb));        // w[63 - r]
            if (b == 0) { t1[a] = wlo; t1[ac] = whi; }              // both partners live in group 0
            else if (s < 32) { t1[a] = wlo; t2[ac] = whi; }
            else { t1[a] = whi; t2[ac] = wlo; }
          }
        }
        dft_odd_p<7, +1>(t1);
#pragma unroll
        for (int a = 0; a < 7; ++a) {
          const f2 one[1] = {t1[a]};
          tmem_st(tcol + 2 * (9 * a + b), one);
        }
        if (b != 0) {
          dft_odd_p<7, +1>(t2);
#pragma unroll
          for (int a = 0; a < 7; ++a) {
            const f2 one[1] = {t2[a]};
            tmem_st(tcol + 2 * (9 * a + bc), one);
          }
        }
      }
      // odd column kq = 0: nine DFT-7 on lanes 0..8 now, seven DFT-9 on lanes 0..6 below
      f2 t7[7];
      if (lane < 9) {
#pragma unroll
        for (int a = 0; a < 7; ++a) {
          const int r = (a * P63::UA + lane * P63::UB) % 63;
          const f2 v = sm->y.y0[r < 32 ? r : 63 - r];
          t7[a] = f2_make(f2_lo(v), (r == 0) ? 0.f : (r < 32 ? f2_hi(v) : -f2_hi(v)));
        }
      }
      simt::tmem_wait_st();
      simt::sync_warp();            // every read of Y is done: the union now holds the correlation row
      if (lane < 9) {
        dft_odd_p<7, +1>(t7);
#pragma unroll
        for (int a = 0; a < 7; ++a) sm->lx[(a * P63::UA + lane * P63::UB) % 63] = t7[a];
      }
      // ---- phase B, pass 2: DFT-9 over b for every a, scattered to natural order
#pragma unroll
      for (int a = 0; a < 7; ++a) {
        f2 t[9];
        {
          f2 v8[8], v1[1];
          tmem_ld(tcol + 18 * a, v8);
          tmem_ld(tcol + 18 * a + 16, v1);
          simt::tmem_wait_ld();
#pragma unroll
          for (int i = 0; i < 8; ++i) t[i] = v8[i];
          t[8] = v1[0];
        }
        dft_odd_p<9, +1>(t);
#pragma unroll
        for (int b = 0; b < 9; ++b) {
          const int kr = P63::out_index(P63::slot(a, b));
          const int c = 65 * kr;
          const float vr = f2_lo(t[b]), vi = f2_hi(t[b]);
          if (kr <= 31) p1[c] = vr;
          else ((b0 + c >= kN4095) ? p1w : p1)[c] = vr;
          if (kr == 0) p2[c] = vi;
          else if (kr >= 32) p2w[c] = vi;
          else ((b0 + 63 * 32 + c >= kN4095) ? p2w : p2)[c] = vi;
          gm = fmaxf(gm, fmaxf(vr, vi));
        }
      }
      simt::sync_warp();
      if (lane < 7) {
        f2 u[9];
#pragma unroll
        for (int b = 0; b < 9; ++b) u[b] = sm->lx[(lane * P63::UA + b * P63::UB) % 63];
        dft_odd_p<9, +1>(u);
#pragma unroll
        for (int b = 0; b < 9; ++b) {
          const int kr = (9 * lane + 7 * b) % 63;
          const float v = f2_lo(u[b]);
          sm->corr[(65 * kr) % kN4095] = v;
          gm = fmaxf(gm, v);
        }
      }
      // ---- fetch the first bins of the NEXT pair while this one is being picked
      const long long next = item + stride;
      const bool has_next = next < n_items;
      if (has_next) {
        pair_it += stride_p;
        frame_it += stride_f;
        if (pair_it >= P) { pair_it -= P; ++frame_it; }
        const long long frame = frame_it;
        const int p = pair_it;
        const int mi = pairs_in_smem ? pairs_s[2 * p] : pairs[2 * p];
        const int mj = pairs_in_smem ? pairs_s[2 * p + 1] : pairs[2 * p + 1];
        si = reinterpret_cast<const f2*>(spec + (frame * M + mi) * kSpecSlots) + lane;
        sj = reinterpret_cast<const f2*>(spec + (frame * M + mj) * kSpecSlots) + lane;
        wb_next = whiten_bound(hq, frame * M + mi, frame * M + mj);
#pragma unroll
        for (int g = 0; g < kLA; ++g)
#pragma unroll
          for (int a5 = 0; a5 < 5; ++a5) { ra[g][a5] = si[P65::slot(a5, g) * 32]; rb[g][a5] = sj[P65::slot(a5, g) * 32]; }
      }
      simt::sync_warp();
      fast_pick_row<WRITE_CORR>(sm->corr, item, gm, lane, pk, wb, k_idx, peak, gmax, flags, corr_out);
      if (!has_next) break;
      item = next;
      wb = wb_next;
      simt::sync_warp();   // the next item overwrites the union
    }
  }
  simt::tmem_fence_before_sync();
  simt::sync_block();
  if (warp == 0) simt::tmem_dealloc512(tbase);
}

}  // namespace pal
