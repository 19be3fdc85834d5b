// pal_dft_small.h -- register-resident odd-length DFT modules (N = 3,5,7,9,13) and the two
// prime-factor composites DFT-63 (7x9) and DFT-65 (5x13) that build the exact length-4095
// transform required by the reference (n = n1+n2-1 = 4095 for 2048-sample frames,
// utils.py:113; SURVEY.md headline fact 2).  All indices are compile-time constants after
// unrolling, so every array below lives in registers.
//
// SIGN = -1: forward  X[k] = sum x[j] exp(-2*pi*i*j*k/N)   (np.fft.fft,  utils.py:114-115)
// SIGN = +1: backward X[k] = sum x[j] exp(+2*pi*i*j*k/N)   (np.fft.ifft without the 1/N, utils.py:118)
#pragma once
#include "pal_simt.h"
#include "pal_dft_consts.h"

namespace pal {

// Symmetric ("half-twiddle") form, valid for any odd N:
//   a_j = x_j + x_{N-j},  b_j = x_j - x_{N-j}
//   X_k, X_{N-k} = x_0 + sum_j a_j cos(2 pi j k / N)  -/+  i * sum_j b_j sin(2 pi j k / N)
// Nearly every operation is an FMA: 36 / 66 / 104 / 204 instructions for N = 5 / 7 / 9 / 13.
template <int N, int SIGN, typename T>
PAL_DEV void dft_odd(T (&xr)[N], T (&xi)[N]) {
  constexpr int H = (N - 1) / 2;
  T ar[H + 1], ai[H + 1], br[H + 1], bi[H + 1];
#pragma unroll
  for (int j = 1; j <= H; ++j) {
    ar[j] = xr[j] + xr[N - j];
    ai[j] = xi[j] + xi[N - j];
    br[j] = xr[j] - xr[N - j];
    bi[j] = xi[j] - xi[N - j];
  }
  const T x0r = xr[0], x0i = xi[0];
  T s0r = x0r, s0i = x0i;
#pragma unroll
  for (int j = 1; j <= H; ++j) {
    s0r += ar[j];
    s0i += ai[j];
  }
  xr[0] = s0r;
  xi[0] = s0i;
#pragma unroll
  for (int k = 1; k <= H; ++k) {
    T cr = x0r, ci = x0i, sr = T(0), si = T(0);
#pragma unroll
    for (int j = 1; j <= H; ++j) {
      const T c = T(TwTab<N>::c((j * k) % N));
      const T s = T(TwTab<N>::s((j * k) % N));
      cr = fma_(ar[j], c, cr);
      ci = fma_(ai[j], c, ci);
      sr = fma_(br[j], s, sr);
      si = fma_(bi[j], s, si);
    }
    if (SIGN < 0) {
      xr[k] = cr + si;      xi[k] = ci - sr;
      xr[N - k] = cr - si;  xi[N - k] = ci + sr;
    } else {
      xr[k] = cr - si;      xi[k] = ci + sr;
      xr[N - k] = cr + si;  xi[N - k] = ci - sr;
    }
  }
}

// Prime-factor composite of two coprime odd lengths A*B held in one thread's registers.
// Slot s (0 <= s < A*B) holds the element whose index is s, i.e. CRT label
// (s mod A, s mod B).  After the call slot s holds the output of index
//     out_index(s) = (B*(s mod A) + A*(s mod B)) mod (A*B)        (Ruritanian map),
// so no twiddle multiplications are needed between the two passes (Good-Thomas).
template <int A, int B> struct Pfa2 {
  static constexpr int N = A * B;
  // CRT basis: UA = 1 (mod A), 0 (mod B); UB = 0 (mod A), 1 (mod B)
  static PAL_HD constexpr int inv_mod(int a, int m) {
    for (int t = 1; t < m; ++t)
      if ((a * t) % m == 1) return t;
    return 0;
  }
  static constexpr int UA = B * inv_mod(B % A, A);
  static constexpr int UB = A * inv_mod(A % B, B);
  static PAL_HD constexpr int slot(int a, int b) { return (a * UA + b * UB) % N; }
  static PAL_HD constexpr int out_index(int s) { return (B * (s % A) + A * (s % B)) % N; }
  // slot that holds output index k after the transform
  static PAL_HD constexpr int slot_of_out(int k) {
    // k = B*ka + A*kb  ->  ka = k * B^{-1} mod A, kb = k * A^{-1} mod B
    return slot((k * inv_mod(B % A, A)) % A, (k * inv_mod(A % B, B)) % B);
  }
};

template <int A, int B, int SIGN, typename T>
PAL_DEV void dft_pfa2(T (&zr)[A * B], T (&zi)[A * B]) {
  using P = Pfa2<A, B>;
  // pass 1: length-A transforms along the first label, one per value of the second label
#pragma unroll
  for (int b = 0; b < B; ++b) {
    T tr[A], ti[A];
#pragma unroll
    for (int a = 0; a < A; ++a) {
      tr[a] = zr[P::slot(a, b)];
      ti[a] = zi[P::slot(a, b)];
    }
    dft_odd<A, SIGN, T>(tr, ti);
#pragma unroll
    for (int a = 0; a < A; ++a) {
      zr[P::slot(a, b)] = tr[a];
      zi[P::slot(a, b)] = ti[a];
    }
  }
  // pass 2: length-B transforms along the second label
#pragma unroll
  for (int a = 0; a < A; ++a) {
    T tr[B], ti[B];
#pragma unroll
    for (int b = 0; b < B; ++b) {
      tr[b] = zr[P::slot(a, b)];
      ti[b] = zi[P::slot(a, b)];
    }
    dft_odd<B, SIGN, T>(tr, ti);
#pragma unroll
    for (int b = 0; b < B; ++b) {
      zr[P::slot(a, b)] = tr[b];
      zi[P::slot(a, b)] = ti[b];
    }
  }
}

// ---- packed (FFMA2) flavour: the same symmetric odd-length module on complex numbers carried as
// (re, im) register pairs.  Every add / multiply-accumulate below is ONE f32x2 instruction
// (half the issue slots of the scalar form); the +-i rotations of the final combine are operand
// modifiers of the FADD2.  Twiddles are broadcast immediates.  N = 5 / 7 / 9 / 13 cost
// 18 / 33 / 52 / 102 instructions.
template <int N, int SIGN>
PAL_DEV void dft_odd_p(f2 (&x)[N]) {
  constexpr int H = (N - 1) / 2;
  f2 a[H + 1], b[H + 1];
#pragma unroll
  for (int j = 1; j <= H; ++j) {
    a[j] = f2_add(x[j], x[N - j]);
    b[j] = f2_sub(x[j], x[N - j]);
  }
  const f2 x0 = x[0];
  f2 s0 = x0;
#pragma unroll
  for (int j = 1; j <= H; ++j) s0 = f2_add(s0, a[j]);
  x[0] = s0;
#pragma unroll
  for (int k = 1; k <= H; ++k) {
    f2 c = x0;
    f2 s = f2_mul(b[1], f2_bcast(float(TwTab<N>::s(k % N))));
#pragma unroll
    for (int j = 1; j <= H; ++j) {
      c = f2_fma(a[j], f2_bcast(float(TwTab<N>::c((j * k) % N))), c);
      if (j > 1) s = f2_fma(b[j], f2_bcast(float(TwTab<N>::s((j * k) % N))), s);
    }
    const f2 is = f2_muli(s);
    if (SIGN < 0) {
      x[k] = f2_sub(c, is);
      x[N - k] = f2_add(c, is);
    } else {
      x[k] = f2_add(c, is);
      x[N - k] = f2_sub(c, is);
    }
  }
}

template <int A, int B, int SIGN>
PAL_DEV void dft_pfa2_p(f2 (&z)[A * B]) {
  using P = Pfa2<A, B>;
#pragma unroll
  for (int b = 0; b < B; ++b) {
    f2 t[A];
#pragma unroll
    for (int a = 0; a < A; ++a) t[a] = z[P::slot(a, b)];
    dft_odd_p<A, SIGN>(t);
#pragma unroll
    for (int a = 0; a < A; ++a) z[P::slot(a, b)] = t[a];
  }
#pragma unroll
  for (int a = 0; a < A; ++a) {
    f2 t[B];
#pragma unroll
    for (int b = 0; b < B; ++b) t[b] = z[P::slot(a, b)];
    dft_odd_p<B, SIGN>(t);
#pragma unroll
    for (int b = 0; b < B; ++b) z[P::slot(a, b)] = t[b];
  }
}

// ---- the length-4095 index algebra shared by every 4095 kernel -------------------------
// 4095 = 63 * 65.  Element e carries CRT label (r, q) = (e mod 63, e mod 65);
//   e(r, q) = (2080 r + 2016 q) mod 4095          (2080 = 1 mod 63, 0 mod 65; 2016 = 0, 1)
// after DFT-65 over q (output label kq) and DFT-63 over r (output label kr) the result of
// index k = (65 kr + 63 kq) mod 4095 is obtained without any twiddle factor.
struct Idx4095 {
  static constexpr int N = 4095, R = 63, Q = 65, UR = 2080, UQ = 2016;
  static PAL_HD constexpr int elem(int r, int q) { return (UR * r + UQ * q) % N; }
  static PAL_HD constexpr int out(int kr, int kq) { return (Q * kr + R * kq) % N; }
};

}  // namespace pal
