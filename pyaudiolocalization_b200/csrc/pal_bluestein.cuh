// pal_bluestein.cuh -- exact DFT of ARBITRARY length n (Bluestein / chirp-z over power-of-two
// FFTs) for the GCC-PHAT path when n = n1+n2-1 is not 4095.
//
// The reference transforms at exactly n = n1+n2-1 points (utils.py:113-118; n = 88199 for the
// README's 1 s @ 44.1 kHz signals, 7999 for 0.25 s @ 16 kHz); padding to a power of two would
// change the PHAT-whitened result (SURVEY.md headline fact 2), so the length-n DFT is evaluated
// exactly as a length-M circular convolution, M = 2^p >= 2n-1:
//     X[k] = w[k] * sum_j (x[j] w[j]) conj(w)[k-j],   w[m] = exp(-+ i pi m^2 / n)
// with m^2 reduced mod 2n in 64-bit integers before the phase is formed (m^2 reaches 7.8e9).
//
// The length-M FFTs are two-pass ("four-step", M = M1 x M2, both <= 1024) radix-2 transforms
// staged in shared memory: a decimation-in-frequency forward pass leaves the spectrum in
// bit-reversed order, the pointwise product with the (equally permuted) chirp spectrum does not
// care, and a decimation-in-time inverse pass restores natural order -- no transposes and no
// bit-reversal copies ever touch global memory.
//   colpass_fwd : load (fused: chirp pre-multiply / PHAT weighting), column FFTs, twiddle
//   rowpass     : row FFT, x chirp spectrum, inverse row FFT, twiddle      (one kernel)
//   colpass_inv : inverse column FFTs, chirp post-multiply, store (fused: spectrum / corr row)
#pragma once
#include "pal_simt.h"
#include "pal_peakpick.cuh"

namespace pal {

struct BluePlan {
  int n;        // DFT length
  int M;        // convolution length (power of two >= 2n-1)
  int M1, M2;   // M = M1 * M2, M1 >= M2
  int lg1, lg2;
};

PAL_HD int ilog2(int v) {
  int l = 0;
  while ((1 << l) < v) ++l;
  return l;
}
PAL_HD BluePlan make_blue_plan(int n) {
  BluePlan p;
  p.n = n;
  int lg = ilog2(2 * n - 1);
  if (lg < 2) lg = 2;
  p.M = 1 << lg;
  p.lg1 = (lg + 1) / 2;
  p.lg2 = lg - p.lg1;
  p.M1 = 1 << p.lg1;
  p.M2 = 1 << p.lg2;
  return p;
}
PAL_DEV unsigned bitrev(unsigned v, int bits) {
#if PAL_GPU
  return __brev(v) >> (32 - bits);
#else
  unsigned r = 0;
  for (int i = 0; i < bits; ++i) r |= ((v >> i) & 1u) << (bits - 1 - i);
  return r;
#endif
}

template <typename T> PAL_DEV cpx<T> cmul(cpx<T> a, cpx<T> b) {
  return cpx<T>{fma_(a.x, b.x, -(a.y * b.y)), fma_(a.x, b.y, a.y * b.x)};
}
template <typename T> PAL_DEV cpx<T> cmulc(cpx<T> a, cpx<T> b) {   // a * conj(b)
  return cpx<T>{fma_(a.x, b.x, a.y * b.y), fma_(a.y, b.x, -(a.x * b.y))};
}

// Device tables of one plan (all in precision T, generated in float64):
//   chirp[m] = exp(-i pi (m^2 mod 2n) / n), m < n        (forward-DFT sign)
//   tw1[k] = exp(-2 pi i k / M1), k < M1/2 ; tw2 likewise ; twM[k] = exp(-2 pi i k / M), k < M
//   bhat[M]  = FFT_M of the wrapped conj(chirp) sequence, in the [k1'][k2'] two-pass layout
template <typename T> struct BlueTables {
  const cpx<T>* chirp;
  const cpx<T>* tw1;
  const cpx<T>* tw2;
  const cpx<T>* twM;
  const cpx<T>* bhat;
};

template <typename T> PAL_DEV void sincospi_(double x, T& s, T& c) {
  double sd, cd;
#if PAL_GPU
  sincospi(x, &sd, &cd);
#else
  sd = std::sin(3.14159265358979323846 * x);
  cd = std::cos(3.14159265358979323846 * x);
#endif
  s = T(sd);
  c = T(cd);
}

// single-precision sin / cos of pi x for an argument already reduced to a few turns
PAL_DEV void sincospif_(float x, float& s, float& c) {
#if PAL_GPU
  sincospif(x, &s, &c);
#else
  s = float(std::sin(3.14159265358979323846 * double(x)));
  c = float(std::cos(3.14159265358979323846 * double(x)));
#endif
}

// grid-stride fill of chirp / twiddle tables
template <typename T>
PAL_DEV void blue_init_tables_body(BluePlan p, cpx<T>* chirp, cpx<T>* tw1, cpx<T>* tw2, cpx<T>* twM) {
  const long long gtid = (long long)simt::bid() * simt::nthreads() + simt::tid();
  const long long gsz = (long long)simt::nblocks() * simt::nthreads();
  for (long long m = gtid; m < p.n; m += gsz) {
    const long long r = (m * m) % (2LL * p.n);
    T s, c;
    sincospi_<T>(double(r) / double(p.n), s, c);
    chirp[m] = cpx<T>{c, -s};
  }
  for (long long k = gtid; k < p.M1 / 2; k += gsz) {
    T s, c;
    sincospi_<T>(2.0 * double(k) / double(p.M1), s, c);
    tw1[k] = cpx<T>{c, -s};
  }
  for (long long k = gtid; k < p.M2 / 2; k += gsz) {
    T s, c;
    sincospi_<T>(2.0 * double(k) / double(p.M2), s, c);
    tw2[k] = cpx<T>{c, -s};
  }
  for (long long k = gtid; k < p.M; k += gsz) {
    T s, c;
    sincospi_<T>(2.0 * double(k) / double(p.M), s, c);
    twM[k] = cpx<T>{c, -s};
  }
}

// ---- radix-2 FFT of a tile in shared memory --------------------------------------------------
// Element (l, c) of the tile sits at l*sl + c*sc.  L = 1 << lg points along l, TC independent
// transforms along c.  Forward: decimation in frequency, natural in -> bit-reversed out.
// Inverse: decimation in time with conjugate twiddles, bit-reversed in -> natural out (unscaled).
// One radix-2 butterfly of either flavour on register values.
template <typename T> PAL_DEV void bfly_dif(T& ur, T& ui, T& vr, T& vi, cpx<T> w) {   // (u, v) -> (u + v, (u - v) w)
  const T sr = ur + vr, si = ui + vi;
  const T dr = ur - vr, di = ui - vi;
  ur = sr; ui = si;
  vr = fma_(dr, w.x, -(di * w.y));
  vi = fma_(dr, w.y, di * w.x);
}
template <typename T> PAL_DEV void bfly_dit(T& ur, T& ui, T& vr, T& vi, cpx<T> w) {   // (u, v) -> (u + v conj(w), u - v conj(w))
  const T tr = fma_(vr, w.x, vi * w.y);
  const T ti = fma_(vi, w.x, -(vr * w.y));
  vr = ur - tr; vi = ui - ti;
  ur = ur + tr; ui = ui + ti;
}

// Consecutive radix-2 stages are fused into one step held in registers: three at a time (radix-8: a third of the
// block barriers, shared-memory round trips and index arithmetic of plain radix-2), then two (radix-4) or one for
// what is left.  Fusing never changes the ORDER of the stages (halves L/2 .. 1 forward, 1 .. L/2 inverse), so the
// bit-reversed hand-over between the forward and the inverse transform is the same for every grouping.
#ifndef PAL_FFT_MAX_FUSE
#define PAL_FFT_MAX_FUSE 3
#endif
template <typename T, int NT>
PAL_DEV void fft_tile(T* re, T* im, int lg, int TC /* power of two */, int sl, int sc, const cpx<T>* tw, bool inverse) {
  const int L = 1 << lg;
  const int lgc = ilog2(TC);
  int st = 0;
  while (st < lg) {
    const int rem = lg - st;
    // radix-8 only where every thread still gets at least two butterflies per step (measured on B200: +8-10 % for
    // 512-point tiles, -5 % for tiles of 128 points and less, where radix-4 keeps more warps busy between barriers)
    const int fuse = (rem >= 3 && PAL_FFT_MAX_FUSE >= 3 && (L >> 3) * TC >= 2 * NT) ? 3 : (rem >= 2 ? 2 : 1);
    if (fuse == 1) {
      const int hl = inverse ? st : (lg - 1 - st);     // log2(half)
      const int half = 1 << hl;
      const int tws = lg - 1 - hl;                     // twiddle stride = L / (2*half)
      const int nb = (L >> 1) * TC;
      for (int t = simt::tid(); t < nb; t += NT) {
        const int c = t & (TC - 1);
        const int b = t >> lgc;
        const int i = b & (half - 1);
        const int p = ((b >> hl) << (hl + 1)) + i;
        const int ap = p * sl + c * sc, aq = ap + half * sl;
        const cpx<T> w = tw[i << tws];
        T ur = re[ap], ui = im[ap], vr = re[aq], vi = im[aq];
        if (!inverse) bfly_dif(ur, ui, vr, vi, w); else bfly_dit(ur, ui, vr, vi, w);
        re[ap] = ur; im[ap] = ui; re[aq] = vr; im[aq] = vi;
      }
    } else if (fuse == 2) {
      // forward: stages with half = H then H/2 (H = 1 << hl); inverse: half = h then 2h (h = 1 << hl)
      const int nq = (L >> 2) * TC;
      if (!inverse) {
        const int hl = lg - 1 - st;                    // first stage: half = 1 << hl
        const int q = 1 << (hl - 1);                   // quarter
        const int tws = lg - 1 - hl;
        for (int t = simt::tid(); t < nq; t += NT) {
          const int c = t & (TC - 1);
          const int b = t >> lgc;
          const int i = b & (q - 1);
          const int p = ((b >> (hl - 1)) << (hl + 1)) + i;
          const int a0 = p * sl + c * sc, a1 = a0 + q * sl, a2 = a0 + 2 * q * sl, a3 = a0 + 3 * q * sl;
          T x0r = re[a0], x0i = im[a0], x1r = re[a1], x1i = im[a1], x2r = re[a2], x2i = im[a2], x3r = re[a3], x3i = im[a3];
          bfly_dif(x0r, x0i, x2r, x2i, tw[i << tws]);
          bfly_dif(x1r, x1i, x3r, x3i, tw[(i + q) << tws]);
          const cpx<T> w2 = tw[i << (tws + 1)];
          bfly_dif(x0r, x0i, x1r, x1i, w2);
          bfly_dif(x2r, x2i, x3r, x3i, w2);
          re[a0] = x0r; im[a0] = x0i; re[a1] = x1r; im[a1] = x1i; re[a2] = x2r; im[a2] = x2i; re[a3] = x3r; im[a3] = x3i;
        }
      } else {
        const int hl = st;                             // first stage: half = h = 1 << hl, second: 2h
        const int h = 1 << hl;
        const int tws = lg - 1 - hl;
        for (int t = simt::tid(); t < nq; t += NT) {
          const int c = t & (TC - 1);
          const int b = t >> lgc;
          const int i = b & (h - 1);
          const int p = ((b >> hl) << (hl + 2)) + i;
          const int a0 = p * sl + c * sc, a1 = a0 + h * sl, a2 = a0 + 2 * h * sl, a3 = a0 + 3 * h * sl;
          T x0r = re[a0], x0i = im[a0], x1r = re[a1], x1i = im[a1], x2r = re[a2], x2i = im[a2], x3r = re[a3], x3i = im[a3];
          const cpx<T> w1 = tw[i << tws];
          bfly_dit(x0r, x0i, x1r, x1i, w1);
          bfly_dit(x2r, x2i, x3r, x3i, w1);
          bfly_dit(x0r, x0i, x2r, x2i, tw[i << (tws - 1)]);
          bfly_dit(x1r, x1i, x3r, x3i, tw[(i + h) << (tws - 1)]);
          re[a0] = x0r; im[a0] = x0i; re[a1] = x1r; im[a1] = x1i; re[a2] = x2r; im[a2] = x2i; re[a3] = x3r; im[a3] = x3i;
        }
      }
    } else {
      // radix-8: eight elements e apart (e = the smallest of the three halves); element k sits at p + k e
      const int n8 = (L >> 3) * TC;
      const int hl = inverse ? st : (lg - 3 - st);     // log2(e)
      const int e = 1 << hl;
      const int tws = lg - 1 - hl;                     // twiddle stride of the stage with half = e
      for (int t = simt::tid(); t < n8; t += NT) {
        const int c = t & (TC - 1);
        const int b = t >> lgc;
        const int i = b & (e - 1);
        const int p = ((b >> hl) << (hl + 3)) + i;
        const int a0 = p * sl + c * sc, da = e * sl;
        T xr[8], xi[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) { xr[k] = re[a0 + k * da]; xi[k] = im[a0 + k * da]; }
        if (!inverse) {
          // halves 4e, 2e, e: position of a butterfly's first element inside its half selects the twiddle
#pragma unroll
          for (int k = 0; k < 4; ++k) bfly_dif(xr[k], xi[k], xr[k + 4], xi[k + 4], tw[(i + k * e) << (tws - 2)]);
          const cpx<T> wb0 = tw[i << (tws - 1)], wb1 = tw[(i + e) << (tws - 1)];
          bfly_dif(xr[0], xi[0], xr[2], xi[2], wb0);
          bfly_dif(xr[1], xi[1], xr[3], xi[3], wb1);
          bfly_dif(xr[4], xi[4], xr[6], xi[6], wb0);
          bfly_dif(xr[5], xi[5], xr[7], xi[7], wb1);
          const cpx<T> wc = tw[i << tws];
#pragma unroll
          for (int k = 0; k < 8; k += 2) bfly_dif(xr[k], xi[k], xr[k + 1], xi[k + 1], wc);
        } else {
          // halves e, 2e, 4e
          const cpx<T> wa = tw[i << tws];
#pragma unroll
          for (int k = 0; k < 8; k += 2) bfly_dit(xr[k], xi[k], xr[k + 1], xi[k + 1], wa);
          const cpx<T> wb0 = tw[i << (tws - 1)], wb1 = tw[(i + e) << (tws - 1)];
          bfly_dit(xr[0], xi[0], xr[2], xi[2], wb0);
          bfly_dit(xr[1], xi[1], xr[3], xi[3], wb1);
          bfly_dit(xr[4], xi[4], xr[6], xi[6], wb0);
          bfly_dit(xr[5], xi[5], xr[7], xi[7], wb1);
#pragma unroll
          for (int k = 0; k < 4; ++k) bfly_dit(xr[k], xi[k], xr[k + 4], xi[k + 4], tw[(i + k * e) << (tws - 2)]);
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) { re[a0 + k * da] = xr[k]; im[a0 + k * da] = xi[k]; }
      }
    }
    st += fuse;
    simt::sync_block();
  }
}

// ---- what is transformed: loaders (time/frequency sample j of transform t) and storers ------
// All loaders return the Bluestein-premultiplied sample a[j] (zero for j >= n).

// the chirp kernel itself: b[m] = conj(chirp[|m|]) wrapped on the M-circle (used once per plan)
template <typename T> struct LoadBhat {
  BluePlan p;
  const cpx<T>* chirp;
  PAL_DEV cpx<T> operator()(long long, int j) const {
    int m = j < p.n ? j : ((p.M - j) < p.n ? p.M - j : -1);
    if (m < 0) return cpx<T>{T(0), T(0)};
    const cpx<T> w = chirp[m];
    return cpx<T>{w.x, -w.y};
  }
};

// ---- two real sequences per complex transform -----------------------------------------------
// Every transform of the GCC-PHAT path has a REAL side: the channels are real (forward) and the correlation
// rows are real (inverse).  Two real sequences therefore share one complex transform, Z = DFT(x1 + i x2):
//     S1[k] = (Z[k] + conj(Z[n-k])) / 2,      S2[k] = (Z[k] - conj(Z[n-k])) / (2i)
// and IDFT(R1 + i R2) = corr1 + i corr2 when R1, R2 are Hermitian.  This halves the number of length-M FFTs in
// both directions.  The spectra are STORED packed (one row Z per channel pair) and unpacked by the PHAT loader,
// which makes S[n-k] == conj(S[k]) hold bit for bit, so R is exactly Hermitian and the two correlation rows
// of a packed inverse transform do not leak into each other beyond the rounding of the transform itself.
template <typename T> PAL_DEV cpx<T> unpack_two_real(const cpx<T>* z, int n, int k, bool second) {
  const cpx<T> a = z[k], b = z[k == 0 ? 0 : n - k];
  if (!second) return cpx<T>{T(0.5) * (a.x + b.x), T(0.5) * (a.y - b.y)};
  return cpx<T>{T(0.5) * (a.y + b.y), T(-0.5) * (a.x - b.x)};
}

// The price of sharing: the rounding residue of the stronger sequence (about 1e-7 of ITS level in float32) lands in the
// weaker one, and PHAT whitening then amplifies it to full scale.  Every row is therefore brought to unit level before
// it is packed, by an exact power of two (no rounding): scales[row] = {2^-e, 2^e}, max|x| 2^-e in [0.5, 1); an
// all-zero row gets {0, 0}, which makes its spectrum exactly zero as in the reference (R = 0, corr = 0) instead of
// the partner's residue.  The loaders undo the scale where absolute levels matter (the 1e-10 of utils.py:117).
template <int NT, typename TS = float>
PAL_DEV void row_scale_body(const TS* sig, long long n_rows, long long ld, int len_even, int len_odd,
                            float* scales /* [n_rows][2] */, char* smem) {
  float* sh = reinterpret_cast<float*>(smem);
  for (long long row = simt::bid(); row < n_rows; row += simt::nblocks()) {
    const int len = (row & 1) ? len_odd : len_even;
    const TS* x = sig + row * ld;
    float mx = 0.f;
    for (int j = simt::tid(); j < len; j += NT) {
      const float v = abs_(float(x[j]));
      // a float64 sample below the float32 range must still count as "not silent" (its row is not all-zero)
      mx = max_(mx, (v == 0.f && x[j] != TS(0)) ? 1.2e-38f : v);
    }
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) mx = max_(mx, simt::shfl_xor(mx, m));
    if (simt::lane() == 0) sh[simt::warp()] = mx;
    simt::sync_block();
    if (simt::tid() == 0) {
      for (int w = 0; w < NT / 32; ++w) mx = max_(mx, sh[w]);
      float sc = 0.f, inv = 0.f;
      if (mx > 0.f && mx < 3.0e38f) {
        int e;
        (void)frexpf(mx, &e);
        e = e < -100 ? -100 : (e > 100 ? 100 : e);
        sc = ldexpf(1.f, -e);
        inv = ldexpf(1.f, e);
      } else if (mx > 0.f) {
        sc = inv = 1.f;      // inf / NaN rows: leave them alone
      }
      scales[2 * row] = sc;
      scales[2 * row + 1] = inv;
    }
    simt::sync_block();
  }
}

// forward DFT of a PAIR of real channels: a[j] = (x1[j] + i x2[j]) * chirp[j]     (np.fft.fft(sig, n), utils.py:114-115)
// frame-major: packed transform g = t_off + t <-> frame g / CP, channels 2c and 2c+1 (c = g % CP, CP = ceil(Mics/2));
// list mode: transform g <-> channel rows row_list[2g], row_list[2g+1].  Rows have `ld` floats; even / odd rows
// hold len_even / len_odd valid samples (n1, n2 of a single unequal pair), zero beyond.
template <typename T, typename TS = float> struct LoadSignal2 {
  BluePlan p;
  const cpx<T>* chirp;
  const TS* sig;           // float32 rows, or float64 rows for float64 callers (then T is double as well)
  long long ld;
  int Mics, CP;
  int len_even, len_odd;
  const int* row_list;
  long long t_off;
  const float* scales;     // [all rows][2] from row_scale_body
  // everything that depends on the transform only is resolved once per work unit (begin), not once per sample
  struct Ctx {
    const TS *xa, *xb;
    int la, lb;
    T sa, sb;
  };
  PAL_DEV Ctx begin(long long t) const {
    const long long g = t_off + t;
    long long ra, rb;
    if (row_list) {
      ra = row_list[2 * g];
      rb = row_list[2 * g + 1];
    } else {
      const long long f = g / CP;
      const int c = int(g - f * CP);
      ra = f * Mics + 2 * c;
      rb = (2 * c + 1 < Mics) ? ra + 1 : -1;
    }
    Ctx x;
    x.xa = sig + ra * ld;
    x.la = (ra & 1) ? len_odd : len_even;
    x.sa = T(scales[2 * ra]);
    x.xb = rb >= 0 ? sig + rb * ld : sig;
    x.lb = rb >= 0 ? ((rb & 1) ? len_odd : len_even) : 0;
    x.sb = rb >= 0 ? T(scales[2 * rb]) : T(0);
    return x;
  }
  // loads are unconditional (indices clamped into the row, values masked): the compiler can then batch the loads of
  // several samples instead of paying one memory round trip per sample
  PAL_DEV cpx<T> operator()(const Ctx& c, int j) const {
    const bool ina = j < c.la, inb = j < c.lb;
    const int ja = ina ? j : 0, jb = inb ? j : 0, jw = j < p.n ? j : 0;
    // the scale is an exact power of two: the product is exact in the sample type
    const T x = T(c.xa[ja] * TS(ina ? c.sa : T(0)));
    const T y = T(c.xb[jb] * TS(inb ? c.sb : T(0)));
    const cpx<T> w = chirp[jw];
    return cpx<T>{fma_(x, w.x, -(y * w.y)), fma_(x, w.y, y * w.x)};
  }
};

// forward DFT of ONE real row (the renderer's base signal)
template <typename T> struct LoadSignal {
  BluePlan p;
  const cpx<T>* chirp;
  const float* sig;
  int len;
  PAL_DEV cpx<T> operator()(long long, int j) const {
    if (j >= len) return cpx<T>{T(0), T(0)};
    const T x = T(sig[j]);
    const cpx<T> w = chirp[j];
    return cpx<T>{x * w.x, x * w.y};
  }
};

// PHAT-weighted cross spectrum of one item at bin k, scaled by 1/n:  R = Si conj(Sj) / (|Si conj(Sj)| + 1e-10)   (utils.py:116-117)
template <typename T> PAL_DEV cpx<T> phat_cross(cpx<T> a, cpx<T> b, T inv_n) {
  const T xr = fma_(a.x, b.x, a.y * b.y);
  const T xi = fma_(a.y, b.x, -(a.x * b.y));
  const T m2 = fma_(xr, xr, xi * xi);
#if PAL_GPU
  // float32: above |R| = 1e-2 the absolute 1e-10 of the reference is far below half an ulp of |R| (6e-10), so
  // |R| + 1e-10 == |R| and the weight is one reciprocal square root (MUFU.RSQ, < 2 ulp) instead of sqrt + divide
  if (sizeof(T) == 4 && m2 > T(1e-4)) {
    const T sc = inv_n * T(rsqrtf(float(m2)));
    return cpx<T>{xr * sc, xi * sc};
  }
#endif
  const T sc = inv_n / (sqrt_(m2) + T(1e-10));
  return cpx<T>{xr * sc, xi * sc};
}

// inverse DFT of TWO PHAT-weighted cross spectra per transform: a[k] = (R_A[k] + i R_B[k]) * conj(chirp[k]) / n,
// items A = t_off + 2t, B = A + 1 of the resident set (B absent when A is the last item).
// frame-major: item it <-> frame it / P, pair it % P, packed spectrum rows frame * CP + mic / 2;
// list mode (rows2): item it owns packed row it (first channel = its i, second = its j).
template <typename T> struct LoadPhat2 {
  BluePlan p;
  const cpx<T>* chirp;
  const cpx<T>* spec;      // packed spectrum rows of the frames (or flagged items) currently resident
  const int* pairs;        // [P][2]
  int Mics, CP, P;
  long long t_off;         // resident item index of item A of transform 0
  long long n_items;       // resident items end here
  bool rows2;
  const float* scales;     // [all rows][2] (row_scale_body), indexed by GLOBAL channel row
  long long frame_base;    // frame-major: global index of resident frame 0
  const int* row_list;     // list mode: global channel rows of every flagged item ...
  long long item_base;     // ... and the list position of resident item 0
  const int* n_items_dev = nullptr;   // when set, the resident item count is read on the device (no host round trip)
  PAL_DEV long long items() const { return n_items_dev ? (long long)*n_items_dev : n_items; }
  // one item: where its two packed spectra live, which half of each it is, and the levels of its two channels
  struct Item {
    const cpx<T>*zi, *zj;
    bool oi, oj, present, dead;
    T ui, uj;
  };
  PAL_DEV Item item(long long it) const {
    Item m;
    m.present = it < items();
    if (!m.present) {
      m.zi = m.zj = spec;
      m.oi = m.oj = false;
      m.dead = true;
      m.ui = m.uj = T(0);
      return m;
    }
    long long gi, gj;
    if (rows2) {
      m.zi = m.zj = spec + it * p.n;
      m.oi = false;
      m.oj = true;
      gi = row_list[2 * (item_base + it)];
      gj = row_list[2 * (item_base + it) + 1];
    } else {
      const long long f = it / P;
      const int pr = int(it - f * P);
      const int mi = pairs[2 * pr], mj = pairs[2 * pr + 1];
      m.zi = spec + (f * CP + (mi >> 1)) * p.n;
      m.zj = spec + (f * CP + (mj >> 1)) * p.n;
      m.oi = mi & 1;
      m.oj = mj & 1;
      gi = (frame_base + f) * Mics + mi;
      gj = (frame_base + f) * Mics + mj;
    }
    m.ui = T(scales[2 * gi + 1]);      // back to the signal's own level (exact power of two)
    m.uj = T(scales[2 * gj + 1]);
    // an item with an all-zero channel: R == 0, the reference's correlation row is exactly zero
    m.dead = scales[2 * gi] == 0.f || scales[2 * gj] == 0.f;
    return m;
  }
  struct Ctx {
    Item a, b;
  };
  PAL_DEV Ctx begin(long long t) const {
    const long long ia = t_off + 2 * t;
    return Ctx{item(ia), item(ia + 1)};
  }
  PAL_DEV cpx<T> one(const Item& m, int k) const {
    cpx<T> a = unpack_two_real<T>(m.zi, p.n, k, m.oi), b = unpack_two_real<T>(m.zj, p.n, k, m.oj);
    a.x *= m.ui; a.y *= m.ui; b.x *= m.uj; b.y *= m.uj;
    return phat_cross<T>(a, b, T(1) / T(p.n));
  }
  PAL_DEV cpx<T> operator()(const Ctx& c, int k) const {
    if (k >= p.n) return cpx<T>{T(0), T(0)};
    const cpx<T> ra = one(c.a, k);
    const cpx<T> rb = c.b.present ? one(c.b, k) : cpx<T>{T(0), T(0)};
    return cmulc(cpx<T>{ra.x - rb.y, ra.y + rb.x}, chirp[k]);
  }
};

// storers receive y[j] = (a conv b)[j] already scaled by 1/M
template <typename T> struct StoreSpectrum {      // S[t][k] = y[k] * chirp[k]
  BluePlan p;
  const cpx<T>* chirp;
  cpx<T>* spec;
  PAL_DEV void operator()(long long t, int k, cpx<T> y) const {
    if (k < p.n) spec[t * p.n + k] = cmul(y, chirp[k]);
  }
};
template <typename T> struct StoreCorr {          // corr[t][k] = Re(y[k] * conj(chirp[k]))
  BluePlan p;
  const cpx<T>* chirp;
  T* corr;
  PAL_DEV void operator()(long long t, int k, cpx<T> y) const {
    if (k < p.n) {
      const cpx<T> w = chirp[k];
      corr[t * p.n + k] = fma_(y.x, w.x, y.y * w.y);
    }
  }
};
// packed inverse: Re -> row 2t, Im -> row 2t+1 of the chunk   (y conj(w) = corr_A + i corr_B)
template <typename T> struct StoreCorr2 {
  BluePlan p;
  const cpx<T>* chirp;
  T* corr;
  long long n_rows;        // corr rows of this launch (the last transform may own a single row)
  LoadPhat2<T> src;        // what was transformed: a dead item's row is written as exact zeros, not as the rounding
                           // residue of the item it shared the transform with
  struct Ctx {
    T *ra, *rb;            // rb == nullptr: no second row
    bool dead_a, dead_b;
  };
  PAL_DEV Ctx begin(long long t) const {
    const long long ia = src.t_off + 2 * t;
    Ctx c;
    c.ra = corr + (2 * t) * p.n;
    c.rb = (2 * t + 1 < (src.n_items_dev ? src.items() : n_rows)) ? c.ra + p.n : nullptr;
    c.dead_a = src.item(ia).dead;
    c.dead_b = src.item(ia + 1).dead;
    return c;
  }
  PAL_DEV void operator()(const Ctx& c, int k, cpx<T> y) const {
    if (k < p.n) {
      const cpx<T> w = chirp[k];
      c.ra[k] = c.dead_a ? T(0) : fma_(y.x, w.x, y.y * w.y);
      if (c.rb) c.rb[k] = c.dead_b ? T(0) : fma_(y.y, w.x, -(y.x * w.y));
    }
  }
};
template <typename T> struct StoreRaw {           // the chirp spectrum itself (plan set-up)
  cpx<T>* out;
  long long M;
  PAL_DEV void operator()(long long t, int j, cpx<T> y) const { out[t * M + j] = y; }
};

// The butterflies' twiddles (L/2 complex numbers) are staged in shared memory once per block: a radix-4 step
// reads three of them per butterfly, and from global memory each read is a long-scoreboard stall.
template <typename T, int NT> PAL_DEV const cpx<T>* stage_twiddles(const cpx<T>* tw, int count, T* smem_after_tile) {
  cpx<T>* s = reinterpret_cast<cpx<T>*>(smem_after_tile);
  for (int i = simt::tid(); i < count; i += NT) s[i] = tw[i];
  simt::sync_block();
  return s;
}
PAL_HD size_t fft_tile_smem(size_t elem_bytes, int L, int lanes) {       // re + im tile, then the L/2 staged twiddles
  return 2 * elem_bytes * size_t(L) * size_t(lanes) + 2 * elem_bytes * size_t(L / 2 > 0 ? L / 2 : 1);
}

// Loaders / storers that need per-transform set-up expose `Ctx begin(t)` and are then called with the context;
// the plain ones are called with the transform index itself.
template <class F> PAL_DEV auto unit_begin(const F& f, long long t, int) -> decltype(f.begin(t)) { return f.begin(t); }
template <class F> PAL_DEV long long unit_begin(const F&, long long t, long) { return t; }

// ---- pass 1: columns forward -------------------------------------------------------------------
// work unit = (transform t, tile of TC adjacent columns j2).  buf[t][r][j2] <- twiddled column FFT
template <typename T, int NT, int TC, class Loader>
PAL_DEV void colpass_fwd_body(BluePlan p, BlueTables<T> tb, Loader load, long long n_tr, cpx<T>* buf, char* smem) {
  const int tc = p.M2 < TC ? p.M2 : TC;
  const int lgt = ilog2(tc);
  const int tiles = p.M2 / tc;
  T* re = reinterpret_cast<T*>(smem);
  T* im = re + p.M1 * tc;
  const cpx<T>* tw1 = stage_twiddles<T, NT>(tb.tw1, p.M1 / 2, im + p.M1 * tc);
  for (long long u = simt::bid(); u < n_tr * tiles; u += simt::nblocks()) {
    const long long t = u / tiles;
    const int j20 = int(u % tiles) * tc;
    const auto ctx = unit_begin(load, t, 0);
    for (int e = simt::tid(); e < p.M1 * tc; e += NT) {
      const int j1 = e >> lgt, c = e & (tc - 1);
      const cpx<T> a = load(ctx, j1 * p.M2 + j20 + c);
      re[e] = a.x;
      im[e] = a.y;
    }
    simt::sync_block();
    fft_tile<T, NT>(re, im, p.lg1, tc, tc, 1, tw1, false);
    cpx<T>* out = buf + t * p.M;
    for (int e = simt::tid(); e < p.M1 * tc; e += NT) {
      const int r = e >> lgt, c = e & (tc - 1);
      const unsigned k1 = bitrev(unsigned(r), p.lg1);
      const int j2 = j20 + c;
      const cpx<T> w = tb.twM[(k1 * unsigned(j2)) & unsigned(p.M - 1)];     // k1 < M1, j2 < M2 <= 1024: 32-bit product
      out[r * p.M2 + j2] = cmul(cpx<T>{re[e], im[e]}, w);
    }
    simt::sync_block();
  }
}

// ---- pass 2: rows: forward FFT, (x chirp spectrum, inverse FFT, conj twiddle) ------------------
// work unit = (transform t, tile of TR adjacent rows r).  The TR rows are contiguous in global memory
// (TR * M2 complex); in shared memory element e of row c sits at e * (TR + 1) + c, i.e. the rows
// are interleaved (and padded) so that the butterflies of one stage touch consecutive banks.
// CONV=false stops after the forward FFT (plan set-up).
template <typename T, int NT, int TR, bool CONV, bool CONJ_BHAT>
PAL_DEV void rowpass_body(BluePlan p, BlueTables<T> tb, long long n_tr, cpx<T>* buf, char* smem) {
  const int tr = p.M1 < TR ? p.M1 : TR;
  const int tiles = p.M1 / tr;
  const int ld = tr + 1;
  T* re = reinterpret_cast<T*>(smem);
  T* im = re + p.M2 * ld;
  const cpx<T>* tw2 = stage_twiddles<T, NT>(tb.tw2, p.M2 / 2, im + p.M2 * ld);
  for (long long u = simt::bid(); u < n_tr * tiles; u += simt::nblocks()) {
    const long long t = u / tiles;
    const int r0 = int(u % tiles) * tr;
    cpx<T>* rows = buf + t * p.M + (long long)r0 * p.M2;
    for (int x = simt::tid(); x < tr * p.M2; x += NT) {
      const int c = x >> p.lg2, e = x & (p.M2 - 1);
      const cpx<T> v = rows[x];
      re[e * ld + c] = v.x;
      im[e * ld + c] = v.y;
    }
    simt::sync_block();
    fft_tile<T, NT>(re, im, p.lg2, tr, ld, 1, tw2, false);
    if (CONV) {
      const cpx<T>* bh = tb.bhat + (long long)r0 * p.M2;
      for (int x = simt::tid(); x < tr * p.M2; x += NT) {
        const int c = x >> p.lg2, e = x & (p.M2 - 1);
        const cpx<T> b = bh[x];
        const cpx<T> a{re[e * ld + c], im[e * ld + c]};
        const cpx<T> v = CONJ_BHAT ? cmulc(a, b) : cmul(a, b);
        re[e * ld + c] = v.x;
        im[e * ld + c] = v.y;
      }
      simt::sync_block();
      fft_tile<T, NT>(re, im, p.lg2, tr, ld, 1, tw2, true);
      for (int x = simt::tid(); x < tr * p.M2; x += NT) {
        const int c = x >> p.lg2, e = x & (p.M2 - 1);
        const unsigned k1 = bitrev(unsigned(r0 + c), p.lg1);
        const cpx<T> w = tb.twM[(k1 * unsigned(e)) & unsigned(p.M - 1)];
        rows[x] = cmulc(cpx<T>{re[e * ld + c], im[e * ld + c]}, w);
      }
    } else {
      for (int x = simt::tid(); x < tr * p.M2; x += NT) {
        const int c = x >> p.lg2, e = x & (p.M2 - 1);
        rows[x] = cpx<T>{re[e * ld + c], im[e * ld + c]};
      }
    }
    simt::sync_block();
  }
}

// ---- pass 3: columns inverse ---------------------------------------------------------------------
template <typename T, int NT, int TC, class Storer>
PAL_DEV void colpass_inv_body(BluePlan p, BlueTables<T> tb, Storer store, long long n_tr, const cpx<T>* buf,
                              char* smem) {
  const int tc = p.M2 < TC ? p.M2 : TC;
  const int lgt = ilog2(tc);
  const int tiles = p.M2 / tc;
  T* re = reinterpret_cast<T*>(smem);
  T* im = re + p.M1 * tc;
  const cpx<T>* tw1 = stage_twiddles<T, NT>(tb.tw1, p.M1 / 2, im + p.M1 * tc);
  const T inv_m = T(1) / T(p.M);
  for (long long u = simt::bid(); u < n_tr * tiles; u += simt::nblocks()) {
    const long long t = u / tiles;
    const int j20 = int(u % tiles) * tc;
    const cpx<T>* in = buf + t * p.M;
    const auto ctx = unit_begin(store, t, 0);
    for (int e = simt::tid(); e < p.M1 * tc; e += NT) {
      const int r = e >> lgt, c = e & (tc - 1);
      const cpx<T> v = in[r * p.M2 + j20 + c];
      re[e] = v.x;
      im[e] = v.y;
    }
    simt::sync_block();
    fft_tile<T, NT>(re, im, p.lg1, tc, tc, 1, tw1, true);
    for (int e = simt::tid(); e < p.M1 * tc; e += NT) {
      const int j1 = e >> lgt, c = e & (tc - 1);
      store(ctx, j1 * p.M2 + j20 + c, cpx<T>{re[e] * inv_m, im[e] * inv_m});
    }
    simt::sync_block();
  }
}

// ---- peak pick over correlation rows in global memory ------------------------------------------
struct RowPickSmem {
  PickScratch ps;
  int out_k[16];
  double out_gmax, out_peak;
};

// near-tie audit of an fp32 decision (SURVEY.md hard part 3).  For num_peaks == 1 the exact answer can differ from
// the float32 one only if (a) another sample INSIDE the window comes within eps of the winner (a different peak may
// be the tallest, or the tallest may have been wrongly killed / spared by a neighbour), (b) a sample within `dist`
// of the winner -- inside the window or not -- comes within eps of it or exceeds it (the distance rule of
// find_peaks may kill the winner, directly or through a chain that starts there), or (c) the winner is within eps
// of the threshold in force.  Samples outside the window and farther than `dist` from the winner cannot matter:
// they are never candidates and can only kill peaks that are lower than the winner.
template <typename T, int NT>
PAL_DEV unsigned tie_audit(const T* c, int n, int c0, int win_half, int dist, int k_best, T h_best, T eps,
                           unsigned pick_flags, T s_abs /* sum |c| of the row, from peakpick_row */, PickScratch* ps) {
  int lo = 0, hi = n - 1;
  const bool windowed = !(pick_flags & PAL_FLAG_FALLBACK_ARGMAX) && win_half >= 0;
  if (windowed) {
    lo = c0 - win_half;
    hi = c0 + win_half;
    lo = lo < 0 ? 0 : lo;
    hi = hi > n - 1 ? n - 1 : hi;
  }
  int cnt = 0;
  for (int k = lo + simt::tid(); k <= hi; k += NT)       // (a); the whole row when there is no window
    if (k != k_best && c[k] >= h_best - eps) ++cnt;
  if (windowed) {                                         // (b): the neighbourhood of the winner that (a) did not cover
    for (int k = k_best - dist + simt::tid(); k <= k_best + dist; k += NT)
      if (k >= 0 && k < n && (k < lo || k > hi) && c[k] >= h_best - eps) ++cnt;
  }
  cnt = block_sum<int, NT>(cnt, ps->isum);
  unsigned fl = 0;
  if (cnt > 0 && s_abs != T(0)) fl |= PAL_FLAG_NEAR_TIE;
  if (!(pick_flags & PAL_FLAG_FALLBACK_ARGMAX) && h_best < s_abs / T(n) + eps) fl |= PAL_FLAG_NEAR_TIE;    // (c)
  return fl;
}

template <typename T, int NT>
PAL_DEV void pick_rows_body(const T* corr, int n, int c0, long long n_rows, const int* item_list, int win_half,
                            int dist, int method, float mult, int num_peaks, float eps, unsigned char* pkmap_ws,
                            int* k_idx, int* k_count, float* peak, float* gmax, unsigned* flags, unsigned extra_flag,
                            unsigned keep_mask, float* corr_out, char* smem) {
  RowPickSmem* sm = reinterpret_cast<RowPickSmem*>(smem);
  unsigned char* pkmap = pkmap_ws + (long long)simt::bid() * ((n + 15) / 16 * 16);
  for (long long rw = simt::bid(); rw < n_rows; rw += simt::nblocks()) {
    const long long item = item_list ? item_list[rw] : rw;
    const T* c = corr + rw * n;
    T g, pk;
    T* og = reinterpret_cast<T*>(&sm->out_gmax);
    T* op = reinterpret_cast<T*>(&sm->out_peak);
    PickResult pr = peakpick_row<T, NT>(c, n, c0, win_half, dist, method, T(mult), num_peaks, pkmap, &sm->ps,
                                        sm->out_k, og, op);
    simt::sync_block();
    g = *og;
    pk = *op;
    const int k0 = sm->out_k[0];
    unsigned fl = pr.flags | extra_flag;
    if (eps > 0.f) fl |= tie_audit<T, NT>(c, n, c0, win_half, dist, k0, pk, T(eps), pr.flags, T(pr.sum_abs), &sm->ps);
    if (simt::tid() == 0) {
      for (int t = 0; t < num_peaks; ++t) k_idx[item * num_peaks + t] = (t < pr.count) ? sm->out_k[t] : -1;
      if (k_count) k_count[item] = pr.count;
      peak[item] = float(pk);
      gmax[item] = float(g);
      flags[item] = (keep_mask ? (flags[item] & keep_mask) : 0u) | fl;
    }
    if (corr_out)
      for (int k = simt::tid(); k < n; k += NT) corr_out[item * n + k] = float(c[k]);
    simt::sync_block();
  }
}

}  // namespace pal
