// pal_fft2.cuh -- the float32 convolution engine of the arbitrary-length (Bluestein) transforms, second generation:
// compile-time plans, packed f32x2 arithmetic, register-blocked radix-8/16/3 steps, fused passes.
//
// A length-n DFT (np.fft.fft / ifft at exactly n = n1+n2-1 points, utils.py:113-118; length 2N in
// signal_processing.py:68-72) is a circular convolution of length M >= 2n-1 with a chirp (pal_bluestein.cuh).  M is no
// longer forced to a power of two: M = M1 x M2 with M1 in {64, 128, 192, 256, 384, 512} (3 * 2^k allowed) and M2 in
// {64, 128, 256, 512}, e.g. 196608 = 384 x 512 for n = 88199 and n = 95999 (262144 before: 25 % fewer points).
//
// One convolution is three kernels ("four-step" FFT, the M1 x M2 matrix stays in global memory / L2 in between):
//   colpass_fwd : loader (chirp pre-multiply / PHAT weighting / Hermitian extension, from pal_bluestein.cuh) straight
//                 into the registers of the first radix step, column FFTs of length M1 in shared memory, last radix
//                 step straight from registers to global memory with the M-point twiddle applied on the way
//   rowpass     : row FFT of length M2, pointwise product with the chirp spectrum and the first inverse radix step fused
//                 in registers (the chirp spectrum is stored in exactly the order the registers hold it), inverse row FFT
//   colpass_inv : conjugate twiddle applied while loading the first inverse radix step, inverse column FFTs, last step
//                 straight from registers into the storer (chirp post-multiply, real part, fade, ...)
// Forward transforms are decimation in frequency (natural order in, digit-reversed out), inverse transforms decimation
// in time (digit-reversed in, natural out), so no permutation pass exists anywhere; the chirp spectrum is produced by
// the same forward kernels and therefore sits in the same digit-reversed order.
//
// Data lives as one complex number per aligned register pair / 8-byte shared-memory word (f2): complex multiplication
// is two packed instructions (FMUL2 + FFMA2 with the swap / negate / broadcast operand modifiers of sm_100), a rotation
// by +-i is free.  A column tile is [M1][TC] with the TC = 16 or 32 columns on consecutive lanes (conflict-free 64-bit
// accesses, 128 / 256-byte global segments); a row tile is [M2][TR + 1] with the TR rows of the tile on consecutive
// lanes (the global rows are transposed on the way in and out; pitch TR + 1 keeps that transposition conflict-free).
// Every index is a compile-time constant or a shift of the thread index.
#pragma once
#include "pal_bluestein.cuh"

namespace pal {
namespace fft2 {

// ---------------------------------------------------------------------------------------------- complex f2 helpers
PAL_DEV f2 cmul(f2 a, f2 w) {    // a * w = a * wr + (i a) * wi
  return f2_fma(f2_muli(a), f2_bcast(f2_hi(w)), f2_mul(a, f2_bcast(f2_lo(w))));
}
PAL_DEV f2 cmulc(f2 a, f2 w) {   // a * conj(w) = a * wr + (-i a) * wi
  return f2_fma(f2_make(f2_hi(a), -f2_lo(a)), f2_bcast(f2_hi(w)), f2_mul(a, f2_bcast(f2_lo(w))));
}
template <bool CONJ> PAL_DEV f2 cmul_t(f2 a, f2 w) { return CONJ ? cmulc(a, w) : cmul(a, w); }
// multiplication by the quarter-turn of the transform direction: -i forward, +i inverse
template <bool INV> PAL_DEV f2 rot90(f2 a) { return INV ? f2_muli(a) : f2_make(f2_hi(a), -f2_lo(a)); }
PAL_DEV f2 as_f2(cpxf a) { return f2_make(a.x, a.y); }
PAL_DEV cpxf as_cpx(f2 a) { return cpxf{f2_lo(a), f2_hi(a)}; }
#if PAL_GPU
PAL_DEV f2 ld_f2(const cpxf* p) { f2 r; r.v = *reinterpret_cast<const unsigned long long*>(p); return r; }
PAL_DEV void st_f2(cpxf* p, f2 v) { *reinterpret_cast<unsigned long long*>(p) = v.v; }
#else
inline f2 ld_f2(const cpxf* p) { return f2{p->x, p->y}; }
inline void st_f2(cpxf* p, f2 v) { p->x = v.lo; p->y = v.hi; }
#endif

// ---------------------------------------------------------------------------------------------- register DFTs
// In-place, natural order in and out: x[k] <- sum_j x[j] exp(-+ 2 pi i j k / R)   (- forward, + when INV)
template <int R, bool INV> struct Dft;
template <bool INV> struct Dft<2, INV> {
  static PAL_DEV void run(f2 (&x)[2]) {
    const f2 s = f2_add(x[0], x[1]);
    x[1] = f2_sub(x[0], x[1]);
    x[0] = s;
  }
};
template <bool INV> struct Dft<3, INV> {
  static PAL_DEV void run(f2 (&x)[3]) {
    const f2 t = f2_add(x[1], x[2]);
    const f2 d = rot90<INV>(f2_sub(x[1], x[2]));                      // -+ i (x1 - x2)
    const f2 m = f2_fma(t, f2_bcast(-0.5f), x[0]);
    x[0] = f2_add(x[0], t);
    x[1] = f2_fma(d, f2_bcast(0.86602540378443864676f), m);
    x[2] = f2_fma(d, f2_bcast(-0.86602540378443864676f), m);
  }
};
template <bool INV> PAL_DEV void dft4(f2& x0, f2& x1, f2& x2, f2& x3) {
  const f2 t0 = f2_add(x0, x2), t1 = f2_sub(x0, x2), t2 = f2_add(x1, x3), t3 = rot90<INV>(f2_sub(x1, x3));
  x0 = f2_add(t0, t2);
  x2 = f2_sub(t0, t2);
  x1 = f2_add(t1, t3);
  x3 = f2_sub(t1, t3);
}
template <bool INV> struct Dft<4, INV> {
  static PAL_DEV void run(f2 (&x)[4]) { dft4<INV>(x[0], x[1], x[2], x[3]); }
};
// a * W8 and a * W8^3 of the transform direction (W8 = exp(-+ i pi / 4))
template <bool INV> PAL_DEV f2 mul_w8_1(f2 a) { return f2_mul(f2_add(a, rot90<INV>(a)), f2_bcast(0.70710678118654752440f)); }
template <bool INV> PAL_DEV f2 mul_w8_3(f2 a) { return f2_mul(f2_sub(rot90<INV>(a), a), f2_bcast(0.70710678118654752440f)); }
template <bool INV> struct Dft<8, INV> {
  static PAL_DEV void run(f2 (&x)[8]) {
    dft4<INV>(x[0], x[2], x[4], x[6]);      // E[k] in x[2k]
    dft4<INV>(x[1], x[3], x[5], x[7]);      // O[k] in x[2k+1]
    const f2 o0 = x[1], o1 = mul_w8_1<INV>(x[3]), o2 = rot90<INV>(x[5]), o3 = mul_w8_3<INV>(x[7]);
    const f2 e0 = x[0], e1 = x[2], e2 = x[4], e3 = x[6];
    x[0] = f2_add(e0, o0); x[4] = f2_sub(e0, o0);
    x[1] = f2_add(e1, o1); x[5] = f2_sub(e1, o1);
    x[2] = f2_add(e2, o2); x[6] = f2_sub(e2, o2);
    x[3] = f2_add(e3, o3); x[7] = f2_sub(e3, o3);
  }
};
// a * W16^e of the transform direction for the generic exponents (1, 3, 9)
template <bool INV> PAL_DEV f2 mul_w16(f2 a, float c, float s) {     // forward twiddle c - i s
  return INV ? cmulc(a, f2_make(c, -s)) : cmul(a, f2_make(c, -s));
}
template <bool INV> struct Dft<16, INV> {
  // j = a + 4 b, k = 4 c + d:  X[4c + d] = sum_a W4^{ac} W16^{ad} sum_b x[a + 4b] W4^{bd}
  static PAL_DEV void run(f2 (&x)[16]) {
#pragma unroll
    for (int a = 0; a < 4; ++a) dft4<INV>(x[a], x[a + 4], x[a + 8], x[a + 12]);       // slot a + 4d = Y_a[d]
    constexpr float c1 = 0.92387953251128675613f, s1 = 0.38268343236508977173f;       // cos / sin (pi / 8)
    x[1 + 4] = mul_w16<INV>(x[1 + 4], c1, s1);              // a d = 1
    x[1 + 8] = mul_w8_1<INV>(x[1 + 8]);                     // 2
    x[1 + 12] = mul_w16<INV>(x[1 + 12], s1, c1);            // 3: cos(3 pi/8) = s1, sin(3 pi/8) = c1
    x[2 + 4] = mul_w8_1<INV>(x[2 + 4]);                     // 2
    x[2 + 8] = rot90<INV>(x[2 + 8]);                        // 4
    x[2 + 12] = mul_w8_3<INV>(x[2 + 12]);                   // 6
    x[3 + 4] = mul_w16<INV>(x[3 + 4], s1, c1);              // 3
    x[3 + 8] = mul_w8_3<INV>(x[3 + 8]);                     // 6
    x[3 + 12] = mul_w16<INV>(x[3 + 12], -c1, -s1);          // 9: cos(9 pi/8) = -c1, sin(9 pi/8) = -s1
#pragma unroll
    for (int d = 0; d < 4; ++d) dft4<INV>(x[4 * d], x[4 * d + 1], x[4 * d + 2], x[4 * d + 3]);   // slot c + 4d = X[4c + d]
    f2 y[16];
#pragma unroll
    for (int c = 0; c < 4; ++c)
#pragma unroll
      for (int d = 0; d < 4; ++d) y[4 * c + d] = x[c + 4 * d];
#pragma unroll
    for (int k = 0; k < 16; ++k) x[k] = y[k];
  }
};

// ---------------------------------------------------------------------------------------------- plans
// radix list of a length-L transform, in the order the forward (decimation in frequency) stages run
template <int L> struct Radices;
template <> struct Radices<64>  { static constexpr int NS = 2; static PAL_HD constexpr int r(int s) { return 8; } };
template <> struct Radices<128> { static constexpr int NS = 2; static PAL_HD constexpr int r(int s) { return s == 0 ? 8 : 16; } };
template <> struct Radices<256> { static constexpr int NS = 2; static PAL_HD constexpr int r(int s) { return 16; } };
template <> struct Radices<512> { static constexpr int NS = 3; static PAL_HD constexpr int r(int s) { return 8; } };
template <> struct Radices<192> { static constexpr int NS = 3; static PAL_HD constexpr int r(int s) { return s == 0 ? 3 : 8; } };
template <> struct Radices<384> { static constexpr int NS = 3; static PAL_HD constexpr int r(int s) { return s == 0 ? 3 : (s == 1 ? 8 : 16); } };
template <int L> PAL_HD constexpr int sub_len(int s) {        // length of the sub-transforms stage s works on
  int len = L;
  for (int i = 0; i < s; ++i) len /= Radices<L>::r(i);
  return len;
}
// frequency index held at position p after the forward transform (digit reversal of the mixed-radix plan)
template <int L> PAL_HD constexpr int freq_of_pos(int p) {
  int k = 0, mult = 1, len = L;
  for (int s = 0; s < Radices<L>::NS; ++s) {
    const int R = Radices<L>::r(s), Q = len / R;
    k += mult * (p / Q);
    p %= Q;
    mult *= R;
    len = Q;
  }
  return k;
}
PAL_HD int freq_of_pos_rt(int L, int p) {
  switch (L) {
    case 64: return freq_of_pos<64>(p);
    case 128: return freq_of_pos<128>(p);
    case 192: return freq_of_pos<192>(p);
    case 256: return freq_of_pos<256>(p);
    case 384: return freq_of_pos<384>(p);
    default: return freq_of_pos<512>(p);
  }
}

// One radix step of TL interleaved length-L transforms.  Element `pos` of the transform on lane `lane` is read through
// src(pos, lane) and written through dst(pos, lane, value): shared-memory tile, loader / storer or global memory.
// `tw` holds exp(-2 pi i x / L), x < L (forward sign).  Forward: butterfly, then twiddle; inverse: conjugate twiddle,
// then butterfly -- the exact transpose, so inverse(forward(x)) = L x for any radix list.
template <int L, int S, bool INV, int LANES, int NT, class Src, class Dst>
PAL_DEV void radix_step(const f2* tw, Src src, Dst dst) {
  constexpr int R = Radices<L>::r(S);
  constexpr int Ls = sub_len<L>(S);
  constexpr int Q = Ls / R;             // distance of the butterfly's elements (a power of two)
  constexpr int NB = L / R;             // butterflies per transform
  constexpr int G = NT / LANES;         // butterflies in flight per block
  static_assert(NT % LANES == 0 && (Q & (Q - 1)) == 0, "plan geometry");
  const int lane = simt::tid() % LANES, grp = simt::tid() / LANES;
#pragma unroll
  for (int it = 0; it < (NB + G - 1) / G; ++it) {
    const int b = grp + it * G;
    if (NB % G != 0 && b >= NB) break;
    const int blk = b / Q, i = b % Q;
    const int p0 = blk * Ls + i;
    f2 x[R];
#pragma unroll
    for (int q = 0; q < R; ++q) x[q] = src(p0 + q * Q, lane);
    if (INV && Q > 1) {
#pragma unroll
      for (int q = 1; q < R; ++q) x[q] = cmulc(x[q], tw[q * i * (L / Ls)]);
    }
    Dft<R, INV>::run(x);
    if (!INV && Q > 1) {
#pragma unroll
      for (int q = 1; q < R; ++q) x[q] = cmul(x[q], tw[q * i * (L / Ls)]);
    }
#pragma unroll
    for (int q = 0; q < R; ++q) dst(p0 + q * Q, lane, x[q]);
  }
}

template <int M1_, int M2_> struct Plan {
  static constexpr int M1 = M1_, M2 = M2_, M = M1_ * M2_;
  static constexpr int TC = (M1_ <= 192) ? 32 : 16;       // columns per column tile (lanes)
  static constexpr int TR = (M2_ <= 128) ? 32 : 16;       // rows per row tile (lanes)
  static constexpr int LDR = TR + 1;                      // row-tile pitch
  static constexpr int RL = Radices<M2_>::r(Radices<M2_>::NS - 1);    // radix of the fused middle step of the row pass
  static_assert(M2_ % TC == 0 && M1_ % TR == 0, "tile geometry");
  static constexpr size_t col_smem = sizeof(f2) * size_t(M1_) * TC + sizeof(f2) * M1_;
  static constexpr size_t row_smem = sizeof(f2) * size_t(M2_) * LDR + sizeof(f2) * M2_;
};

// tables of one (plan, n):  chirp[n] (pal_bluestein.cuh), tw1[M1], tw2[M2] (stage twiddles, forward sign),
// twf[M] = exp(-2 pi i k1(r) j2 / M) in buffer order [r][j2], bhat[M] = chirp spectrum / M in register order
struct Tables {
  const cpxf* chirp;
  const cpxf* tw1;
  const cpxf* tw2;
  const cpxf* twf;
  const cpxf* bhat;
};

// grid-stride fill of chirp and twiddle tables (float64 phases, rounded once)
PAL_DEV void init_tables_body(int n, int M1, int M2, cpxf* chirp, cpxf* tw1, cpxf* tw2, cpxf* twf) {
  const long long gtid = (long long)simt::bid() * simt::nthreads() + simt::tid();
  const long long gsz = (long long)simt::nblocks() * simt::nthreads();
  const long long M = (long long)M1 * M2;
  for (long long m = gtid; m < n; m += gsz) {
    const long long r = (m * m) % (2LL * n);
    float s, c;
    sincospi_<float>(double(r) / double(n), s, c);
    chirp[m] = cpxf{c, -s};
  }
  for (long long k = gtid; k < M1; k += gsz) {
    float s, c;
    sincospi_<float>(2.0 * double(k) / double(M1), s, c);
    tw1[k] = cpxf{c, -s};
  }
  for (long long k = gtid; k < M2; k += gsz) {
    float s, c;
    sincospi_<float>(2.0 * double(k) / double(M2), s, c);
    tw2[k] = cpxf{c, -s};
  }
  for (long long e = gtid; e < M; e += gsz) {
    const int r = int(e / M2), j2 = int(e % M2);
    float s, c;
    sincospi_<float>(2.0 * double((long long)freq_of_pos_rt(M1, r) * j2) / double(M), s, c);      // k1 j2 < M: no reduction needed
    twf[e] = cpxf{c, -s};
  }
}

template <int NT> PAL_DEV const f2* stage_tw(const cpxf* g, int count, f2* s) {
  for (int i = simt::tid(); i < count; i += NT) s[i] = ld_f2(g + i);
  simt::sync_block();
  return s;
}

// ---------------------------------------------------------------------------------------------- pass 1: columns forward
// work unit = (transform t, tile of TC adjacent columns).  buf[t][r][j2] <- twiddled column FFT (r = digit-reversed k1)
// one work unit: transform t, column tile tile_i (ends with a block barrier: the tile is free again)
template <class P, int NT, class Loader>
PAL_DEV void colpass_fwd_unit(const Tables& tb, const Loader& load, long long t, int tile_i, cpxf* buf, f2* tile, const f2* tw) {
  constexpr int L = P::M1, TC = P::TC, NS = Radices<L>::NS;
  auto rd = [&](int pos, int lane) { return tile[pos * TC + lane]; };
  auto wr = [&](int pos, int lane, f2 v) { tile[pos * TC + lane] = v; };
  const int j20 = tile_i * TC;
  const auto ctx = unit_begin(load, t, 0);
  radix_step<L, 0, false, TC, NT>(tw, [&](int pos, int lane) { return as_f2(load(ctx, pos * P::M2 + j20 + lane)); }, wr);
  simt::sync_block();
  if (NS == 3) {
    radix_step<L, NS == 3 ? 1 : 0, false, TC, NT>(tw, rd, wr);
    simt::sync_block();
  }
  cpxf* out = buf + t * P::M;
  radix_step<L, NS - 1, false, TC, NT>(tw, rd, [&](int pos, int lane, f2 v) {
    const int idx = pos * P::M2 + j20 + lane;
    st_f2(out + idx, cmul(v, ld_f2(tb.twf + idx)));
  });
  simt::sync_block();
}
template <class P, int NT, class Loader>
PAL_DEV void colpass_fwd_body(Tables tb, Loader load, long long n_tr, cpxf* buf, char* smem) {
  constexpr int L = P::M1, TC = P::TC, tiles = P::M2 / TC;
  f2* tile = reinterpret_cast<f2*>(smem);
  const f2* tw = stage_tw<NT>(tb.tw1, L, tile + L * TC);
  for (long long u = simt::bid(); u < n_tr * tiles; u += simt::nblocks())
    colpass_fwd_unit<P, NT, Loader>(tb, load, u / tiles, int(u % tiles), buf, tile, tw);
}

// ---------------------------------------------------------------------------------------------- pass 2: rows
// work unit = (transform t, tile of TR adjacent rows; TR * M2 contiguous complex numbers in global memory).
// MODE 0: x bhat, 1: x conj(bhat) (inverse-direction Bluestein), 2: plan set-up -- stop after the forward FFT and write
// the chirp spectrum, scaled by 1/M, in the register order of the fused middle step:
//   bhat[((row_tile * (M2 / RL) + butterfly) * TR + lane) * RL + q]
template <class P, int NT, int MODE>
PAL_DEV void rowpass_unit(const Tables& tb, long long t, int rt, cpxf* buf, cpxf* bhat_out, f2* tile, const f2* tw) {
  constexpr int L = P::M2, TR = P::TR, LD = P::LDR, NS = Radices<L>::NS, RL = P::RL;
  constexpr int NB = L / RL, G = NT / TR;
  auto rd = [&](int pos, int lane) { return tile[pos * LD + lane]; };
  auto wr = [&](int pos, int lane, f2 v) { tile[pos * LD + lane] = v; };
  const int lane = simt::tid() % TR, grp = simt::tid() / TR;
  cpxf* rows = buf + t * P::M + (long long)rt * TR * L;
  for (int x = simt::tid(); x < TR * L; x += NT) tile[(x % L) * LD + x / L] = ld_f2(rows + x);       // transpose in
  simt::sync_block();
  radix_step<L, 0, false, TR, NT>(tw, rd, wr);
  simt::sync_block();
  if (NS == 3) {
    radix_step<L, NS == 3 ? 1 : 0, false, TR, NT>(tw, rd, wr);
    simt::sync_block();
  }
  // fused middle: last forward step (consecutive elements, no twiddle), x chirp spectrum, first inverse step
#pragma unroll
  for (int it = 0; it < (NB + G - 1) / G; ++it) {
    const int b = grp + it * G;
    if (NB % G != 0 && b >= NB) break;
    f2 x[RL];
#pragma unroll
    for (int q = 0; q < RL; ++q) x[q] = tile[(b * RL + q) * LD + lane];
    Dft<RL, false>::run(x);
    const size_t bo = ((size_t(rt) * NB + b) * TR + lane) * RL;
    if (MODE == 2) {
      const f2 sc = f2_bcast(1.0f / float(P::M));
#pragma unroll
      for (int q = 0; q < RL; ++q) st_f2(bhat_out + bo + q, f2_mul(x[q], sc));
    } else {
      const float4* bp = reinterpret_cast<const float4*>(tb.bhat + bo);
#pragma unroll
      for (int q = 0; q < RL; q += 2) {
        const float4 w = bp[q / 2];
        x[q] = cmul_t<MODE == 1>(x[q], f2_make(w.x, w.y));
        x[q + 1] = cmul_t<MODE == 1>(x[q + 1], f2_make(w.z, w.w));
      }
      Dft<RL, true>::run(x);
#pragma unroll
      for (int q = 0; q < RL; ++q) tile[(b * RL + q) * LD + lane] = x[q];
    }
  }
  simt::sync_block();
  if (MODE != 2) {
    if (NS == 3) {
      radix_step<L, NS == 3 ? 1 : 0, true, TR, NT>(tw, rd, wr);
      simt::sync_block();
    }
    radix_step<L, 0, true, TR, NT>(tw, rd, wr);
    simt::sync_block();
    for (int x = simt::tid(); x < TR * L; x += NT) st_f2(rows + x, tile[(x % L) * LD + x / L]);         // transpose out
    simt::sync_block();
  }
}
template <class P, int NT, int MODE>
PAL_DEV void rowpass_body(Tables tb, long long n_tr, cpxf* buf, cpxf* bhat_out, char* smem) {
  constexpr int L = P::M2, tiles = P::M1 / P::TR;
  f2* tile = reinterpret_cast<f2*>(smem);
  const f2* tw = stage_tw<NT>(tb.tw2, L, tile + L * P::LDR);
  for (long long u = simt::bid(); u < n_tr * tiles; u += simt::nblocks())
    rowpass_unit<P, NT, MODE>(tb, u / tiles, int(u % tiles), buf, bhat_out, tile, tw);
}

// a storer may reduce something over the whole work unit (e.g. the row maximum): finish(ctx, t, tile, scratch) is then
// called by every thread of the block once the unit's samples have been stored; `scratch` is free shared memory
template <class S, class C>
PAL_DEV auto unit_end(const S& s, C& c, long long t, int tile, float* scratch, int) -> decltype(s.finish(c, t, tile, scratch), void()) {
  s.finish(c, t, tile, scratch);
}
template <class S, class C> PAL_DEV void unit_end(const S&, C&, long long, int, float*, long) {}

// ---------------------------------------------------------------------------------------------- pass 3: columns inverse
template <class P, int NT, class Storer>
PAL_DEV void colpass_inv_unit(const Tables& tb, const Storer& store, long long t, int tile_i, const cpxf* buf, f2* tile, const f2* tw) {
  constexpr int L = P::M1, TC = P::TC, NS = Radices<L>::NS;
  auto rd = [&](int pos, int lane) { return tile[pos * TC + lane]; };
  auto wr = [&](int pos, int lane, f2 v) { tile[pos * TC + lane] = v; };
  const int j20 = tile_i * TC;
  const cpxf* in = buf + t * P::M;
  auto ctx = unit_begin(store, t, 0);
  radix_step<L, NS - 1, true, TC, NT>(tw, [&](int pos, int lane) {
    const int idx = pos * P::M2 + j20 + lane;
    return cmulc(ld_f2(in + idx), ld_f2(tb.twf + idx));
  }, wr);
  simt::sync_block();
  if (NS == 3) {
    radix_step<L, NS == 3 ? 1 : 0, true, TC, NT>(tw, rd, wr);
    simt::sync_block();
  }
  radix_step<L, 0, true, TC, NT>(tw, rd, [&](int pos, int lane, f2 v) { store(ctx, pos * P::M2 + j20 + lane, as_cpx(v)); });
  simt::sync_block();
  unit_end(store, ctx, t, tile_i, reinterpret_cast<float*>(tile), 0);
}
template <class P, int NT, class Storer>
PAL_DEV void colpass_inv_body(Tables tb, Storer store, long long n_tr, const cpxf* buf, char* smem) {
  constexpr int L = P::M1, TC = P::TC, tiles = P::M2 / TC;
  f2* tile = reinterpret_cast<f2*>(smem);
  const f2* tw = stage_tw<NT>(tb.tw1, L, tile + L * TC);
  for (long long u = simt::bid(); u < n_tr * tiles; u += simt::nblocks())
    colpass_inv_unit<P, NT, Storer>(tb, store, u / tiles, int(u % tiles), buf, tile, tw);
}

// ---------------------------------------------------------------------------------------------- plan selection
// supported convolution lengths, ascending: {M1, M2}
struct PlanDims { int M1, M2; };
constexpr int kNumPlans = 9;
PAL_HD constexpr PlanDims plan_dims(int i) {
  constexpr PlanDims d[kNumPlans] = {{64, 64}, {128, 64}, {128, 128}, {192, 128}, {256, 128}, {256, 256}, {512, 256},
                                     {384, 512}, {512, 512}};
  return d[i];
}
// index of the smallest plan with M >= 2n - 1, or -1 (the caller falls back to the first-generation engine)
inline int choose_plan(int n) {
  if (n < 1025) return -1;          // tiny transforms: a 64 x 64 convolution would be mostly padding
  for (int i = 0; i < kNumPlans; ++i)
    if ((long long)plan_dims(i).M1 * plan_dims(i).M2 >= 2LL * n - 1) return i;
  return -1;
}

// host-side dispatch: f(Plan<M1, M2>{}) for plan index `id`
template <class F> inline void with_plan(int id, F&& f) {
  switch (id) {
    case 0: f(Plan<64, 64>{}); break;
    case 1: f(Plan<128, 64>{}); break;
    case 2: f(Plan<128, 128>{}); break;
    case 3: f(Plan<192, 128>{}); break;
    case 4: f(Plan<256, 128>{}); break;
    case 5: f(Plan<256, 256>{}); break;
    case 6: f(Plan<512, 256>{}); break;
    case 7: f(Plan<384, 512>{}); break;
    default: f(Plan<512, 512>{}); break;
  }
}

}  // namespace fft2
}  // namespace pal

namespace pal {
// ---------------------------------------------------------------------------------------------- whole convolution in one CTA
// For M <= 16384 the M1 x M2 matrix fits one CTA's shared memory (128 x 129 x 8 B = 129 KB with the pitch that keeps both
// the column passes -- lanes along a row -- and the row passes -- lanes along a column -- free of bank conflicts).  The
// three kernels of the general engine then collapse into one: loader -> column FFT -> twiddle -> row FFT -> x chirp
// spectrum -> inverse row FFT -> conjugate twiddle -> inverse column FFT -> storer, 12 shared-memory accesses per point
// instead of 20 plus three round trips through global memory.  Same radix steps, same digit-reversed orders, same twf
// table; the chirp spectrum is kept in the order THIS kernel's middle step holds it (bhat_s[pos * M1 + row], built by
// the kernel itself in MODE 2).
namespace fft2 {

template <class P> struct SmemConv {
  static constexpr int PITCH = P::M2 + 1;
  static constexpr size_t smem = sizeof(f2) * size_t(P::M1) * PITCH + sizeof(f2) * (P::M1 + P::M2);
  static constexpr bool fits = P::M <= 16384;       // power-of-two plans 64 x 64, 128 x 64, 128 x 128
};

// MODE 0: x bhat_s, 1: x conj(bhat_s), 2: plan set-up (forward only, bhat_s <- spectrum / M)
template <class P, int NT, int MODE, class Loader, class Storer>
PAL_DEV void conv_smem_body(Tables tb, const cpxf* bhat_s, cpxf* bhat_out, Loader load, Storer store, long long n_tr, char* smem) {
  constexpr int M1 = P::M1, M2 = P::M2, PITCH = SmemConv<P>::PITCH;
  constexpr int NS1 = Radices<M1>::NS, NS2 = Radices<M2>::NS, RL = P::RL;
  f2* A = reinterpret_cast<f2*>(smem);
  f2* tw1 = A + M1 * PITCH;
  f2* tw2 = tw1 + M1;
  for (int i = simt::tid(); i < M1; i += NT) tw1[i] = ld_f2(tb.tw1 + i);
  for (int i = simt::tid(); i < M2; i += NT) tw2[i] = ld_f2(tb.tw2 + i);
  simt::sync_block();
  // column passes: element `pos` (row index) of column `lane`;  row passes: element `pos` (column index) of row `lane`
  auto crd = [&](int pos, int lane) { return A[pos * PITCH + lane]; };
  auto cwr = [&](int pos, int lane, f2 v) { A[pos * PITCH + lane] = v; };
  auto rrd = [&](int pos, int lane) { return A[lane * PITCH + pos]; };
  auto rwr = [&](int pos, int lane, f2 v) { A[lane * PITCH + pos] = v; };
  for (long long t = simt::bid(); t < n_tr; t += simt::nblocks()) {
    // ---- forward columns: loader -> registers -> shared memory; last step multiplies the M-point twiddle
    {
      const auto ctx = unit_begin(load, t, 0);
      radix_step<M1, 0, false, M2, NT>(tw1, [&](int pos, int lane) { return as_f2(load(ctx, pos * M2 + lane)); }, cwr);
    }
    simt::sync_block();
    if (NS1 == 3) {
      radix_step<M1, NS1 == 3 ? 1 : 0, false, M2, NT>(tw1, crd, cwr);
      simt::sync_block();
    }
    radix_step<M1, NS1 - 1, false, M2, NT>(tw1, crd, [&](int pos, int lane, f2 v) {
      A[pos * PITCH + lane] = cmul(v, ld_f2(tb.twf + pos * M2 + lane));
    });
    simt::sync_block();
    // ---- forward rows
    radix_step<M2, 0, false, M1, NT>(tw2, rrd, rwr);
    simt::sync_block();
    if (NS2 == 3) {
      radix_step<M2, NS2 == 3 ? 1 : 0, false, M1, NT>(tw2, rrd, rwr);
      simt::sync_block();
    }
    // ---- middle: last forward row step, x chirp spectrum, first inverse row step, all in registers
    {
      constexpr int NB = M2 / RL, G = NT / M1;
      const int lane = simt::tid() % M1, grp = simt::tid() / M1;
#pragma unroll
      for (int it = 0; it < (NB + G - 1) / G; ++it) {
        const int b = grp + it * G;
        if (NB % G != 0 && b >= NB) break;
        f2 x[RL];
#pragma unroll
        for (int q = 0; q < RL; ++q) x[q] = A[lane * PITCH + b * RL + q];
        Dft<RL, false>::run(x);
        if (MODE == 2) {
          const f2 sc = f2_bcast(1.0f / float(P::M));
#pragma unroll
          for (int q = 0; q < RL; ++q) st_f2(bhat_out + (b * RL + q) * M1 + lane, f2_mul(x[q], sc));
        } else {
#pragma unroll
          for (int q = 0; q < RL; ++q) x[q] = cmul_t<MODE == 1>(x[q], ld_f2(bhat_s + (b * RL + q) * M1 + lane));
          Dft<RL, true>::run(x);
#pragma unroll
          for (int q = 0; q < RL; ++q) A[lane * PITCH + b * RL + q] = x[q];
        }
      }
    }
    simt::sync_block();
    if (MODE == 2) continue;
    // ---- inverse rows
    if (NS2 == 3) {
      radix_step<M2, NS2 == 3 ? 1 : 0, true, M1, NT>(tw2, rrd, rwr);
      simt::sync_block();
    }
    radix_step<M2, 0, true, M1, NT>(tw2, rrd, rwr);
    simt::sync_block();
    // ---- inverse columns: conjugate twiddle on the way in, storer on the way out
    radix_step<M1, NS1 - 1, true, M2, NT>(tw1, [&](int pos, int lane) {
      return cmulc(A[pos * PITCH + lane], ld_f2(tb.twf + pos * M2 + lane));
    }, cwr);
    simt::sync_block();
    if (NS1 == 3) {
      radix_step<M1, NS1 == 3 ? 1 : 0, true, M2, NT>(tw1, crd, cwr);
      simt::sync_block();
    }
    {
      auto ctx = unit_begin(store, t, 0);
      radix_step<M1, 0, true, M2, NT>(tw1, crd, [&](int pos, int lane, f2 v) { store(ctx, pos * M2 + lane, as_cpx(v)); });
      simt::sync_block();
      unit_end(store, ctx, t, 0, reinterpret_cast<float*>(A), 0);
    }
  }
}

}  // namespace fft2
}  // namespace pal
