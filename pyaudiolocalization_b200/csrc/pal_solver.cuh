// pal_solver.cuh -- batched position solve for scene sweeps.
//
// The reference solves one scene at a time on the host: residuals r_p = (|x - m_j| - |x - m_i|) - c * td_p over the
// microphone pairs (utils.py:384-405), minimised by scipy's bounded least_squares from a few starting points inside a
// box derived from the array extent and the 75th percentile of c |td| (utils.py:364-382, main.py:246-274).  That stays
// the path of `localize_sound_source`.  A sweep over a million scenes cannot afford a Python call per scene, so this
// kernel runs the same least-squares problem -- same residuals, same weights, same box -- for every scene of a batch:
// one warp per scene, float64, Levenberg-Marquardt on the 3 x 3 normal equations with an active-set treatment of the box.
// At an interior minimum it agrees with scipy.optimize.least_squares run to tight tolerances to < 1e-6 m
// (tests/test_gpu_solver.py); the reference's own stopping rule (ftol = xtol = gtol = 1e-6, relative) stops earlier
// than that, so the two can differ by what scipy leaves on the table.
#pragma once
#include "pal_simt.h"

namespace pal {

struct SolveParams {
  int n_mics, n_pairs, max_iter;
  double c, buffer;         // speed of sound; box margin of dynamic_bounds_extended (5.0 in main.py:246)
  double xtol, ftol, gtol;  // stop: |step| <= xtol (xtol + |x|), relative cost decrease <= ftol, |J^T r|_inf <= gtol
};

PAL_DEV double wsum_d(double v) {
#pragma unroll
  for (int m = 16; m >= 1; m >>= 1) v += simt::shfl_xor(v, m);
  return v;
}
PAL_DEV double wmax_d(double v) {
#pragma unroll
  for (int m = 16; m >= 1; m >>= 1) {
    const double o = simt::shfl_xor(v, m);
    v = o > v ? o : v;
  }
  return v;
}
PAL_DEV double wmin_d(double v) {
#pragma unroll
  for (int m = 16; m >= 1; m >>= 1) {
    const double o = simt::shfl_xor(v, m);
    v = o < v ? o : v;
  }
  return v;
}

// cost 0.5 sum (w r)^2, gradient g = J^T (w r) and A = J^T J (6 unique entries) at x; all lanes get the sums
PAL_DEV double solve_eval(const double* mic, const int* pairs, const double* td, const double* w, int P, double c,
                          const double x[3], double g[3], double A[6], bool need_jac) {
  double cost = 0.0, g0 = 0.0, g1 = 0.0, g2 = 0.0, a00 = 0.0, a01 = 0.0, a02 = 0.0, a11 = 0.0, a12 = 0.0, a22 = 0.0;
  for (int p = simt::lane(); p < P; p += 32) {
    const double* mi = mic + 3 * pairs[2 * p];
    const double* mj = mic + 3 * pairs[2 * p + 1];
    const double ix = x[0] - mi[0], iy = x[1] - mi[1], iz = x[2] - mi[2];
    const double jx = x[0] - mj[0], jy = x[1] - mj[1], jz = x[2] - mj[2];
    const double di = sqrt(ix * ix + iy * iy + iz * iz), dj = sqrt(jx * jx + jy * jy + jz * jz);
    const double wp = w ? w[p] : 1.0;
    const double r = ((dj - di) - c * td[p]) * wp;
    cost += r * r;
    if (need_jac) {
      const double ii = di > 0.0 ? 1.0 / di : 0.0, ij = dj > 0.0 ? 1.0 / dj : 0.0;
      const double j0 = (jx * ij - ix * ii) * wp, j1 = (jy * ij - iy * ii) * wp, j2 = (jz * ij - iz * ii) * wp;
      g0 += j0 * r; g1 += j1 * r; g2 += j2 * r;
      a00 += j0 * j0; a01 += j0 * j1; a02 += j0 * j2; a11 += j1 * j1; a12 += j1 * j2; a22 += j2 * j2;
    }
  }
  cost = 0.5 * wsum_d(cost);
  if (need_jac) {
    g[0] = wsum_d(g0); g[1] = wsum_d(g1); g[2] = wsum_d(g2);
    A[0] = wsum_d(a00); A[1] = wsum_d(a01); A[2] = wsum_d(a02); A[3] = wsum_d(a11); A[4] = wsum_d(a12); A[5] = wsum_d(a22);
  }
  return cost;
}

// solve (A + lam diag(A)) d = -g for the symmetric 3 x 3 A = [a00 a01 a02; . a11 a12; . . a22]; false if singular
PAL_DEV bool solve3(const double A[6], const double g[3], double lam, double d[3]) {
  const double a00 = A[0] * (1.0 + lam) + 1e-300, a11 = A[3] * (1.0 + lam) + 1e-300, a22 = A[5] * (1.0 + lam) + 1e-300;
  const double a01 = A[1], a02 = A[2], a12 = A[4];
  const double c00 = a11 * a22 - a12 * a12, c01 = a02 * a12 - a01 * a22, c02 = a01 * a12 - a02 * a11;
  const double det = a00 * c00 + a01 * c01 + a02 * c02;
  if (!(fabs(det) > 0.0) || !(det == det)) return false;
  const double c11 = a00 * a22 - a02 * a02, c12 = a01 * a02 - a00 * a12, c22 = a00 * a11 - a01 * a01;
  const double inv = -1.0 / det;
  d[0] = (c00 * g[0] + c01 * g[1] + c02 * g[2]) * inv;
  d[1] = (c01 * g[0] + c11 * g[1] + c12 * g[2]) * inv;
  d[2] = (c02 * g[0] + c12 * g[1] + c22 * g[2]) * inv;
  return true;
}

// One warp per scene (grid-stride).  mics [S or 1][n_mics][3] (mic_stride 0: shared array), pairs [P][2],
// tdoa [S][P] seconds, weights [P] or nullptr, x0 [S][3] or nullptr (start at the array centroid),
// lo / hi [S][3] or nullptr (the box of utils.py:364-382 is then formed here from the scene's own mics and TDOAs;
// `pct_scratch` [warps in flight][P] holds the c |td| values while their 75th percentile is selected).
// out_pos [S][3], out_cost [S] (0.5 sum r^2, scipy's `cost`), out_iter [S] (iterations; negative: did not converge).
template <int NT>
PAL_DEV void solve_positions_body(SolveParams sp, const double* mics, long long mic_stride, const int* pairs, const double* tdoa,
                                  const double* weights, const double* x0, const double* lo_in, const double* hi_in,
                                  long long n_scenes, double* pct_scratch, double* out_pos, double* out_cost, int* out_iter) {
  const int lane = simt::lane();
  const int wpb = NT / 32;
  const long long warp_global = (long long)simt::bid() * wpb + simt::warp();
  double* tmp = pct_scratch + warp_global * sp.n_pairs;
  for (long long s = warp_global; s < n_scenes; s += (long long)simt::nblocks() * wpb) {
    const double* mic = mics + s * mic_stride;
    const double* td = tdoa + s * sp.n_pairs;
    const int P = sp.n_pairs;
    // ---- box (dynamic_bounds_extended): mic extent +- (buffer + max(percentile75(c |td|), 1))
    double lo[3], hi[3];
    if (lo_in && hi_in) {
      for (int a = 0; a < 3; ++a) { lo[a] = lo_in[3 * s + a]; hi[a] = hi_in[3 * s + a]; }
    } else {
      for (int p = lane; p < P; p += 32) tmp[p] = sp.c * fabs(td[p]);
      simt::sync_warp();
      // numpy's percentile (linear): position 0.75 (P - 1) between the order statistics; rank by counting (P is small)
      const double pos = 0.75 * double(P - 1);
      const int k0 = int(pos), k1 = k0 + 1 < P ? k0 + 1 : k0;
      double v0 = 0.0, v1 = 0.0;
      for (int p = lane; p < P; p += 32) {
        const double v = tmp[p];
        int less = 0, eq = 0;
        for (int q = 0; q < P; ++q) { less += tmp[q] < v; eq += (tmp[q] == v && q < p); }
        const int rank = less + eq;          // a permutation of 0 .. P-1
        if (rank == k0) v0 = v;
        if (rank == k1) v1 = v;
      }
      v0 = wsum_d(v0);
      v1 = wsum_d(v1);
      double margin = P > 0 ? v0 + (v1 - v0) * (pos - double(k0)) : 0.0;
      margin = P > 0 ? (margin > 1.0 ? margin : 1.0) : 0.0;
      margin += sp.buffer;
      for (int a = 0; a < 3; ++a) {
        double mn = 1e300, mx = -1e300;
        for (int m = lane; m < sp.n_mics; m += 32) { mn = fmin(mn, mic[3 * m + a]); mx = fmax(mx, mic[3 * m + a]); }
        lo[a] = wmin_d(mn) - margin;
        hi[a] = wmax_d(mx) + margin;
      }
      simt::sync_warp();
    }
    double x[3];
    if (x0) {
      for (int a = 0; a < 3; ++a) x[a] = x0[3 * s + a];
    } else {
      for (int a = 0; a < 3; ++a) {
        double sum = 0.0;
        for (int m = lane; m < sp.n_mics; m += 32) sum += mic[3 * m + a];
        x[a] = wsum_d(sum) / double(sp.n_mics);
      }
    }
    for (int a = 0; a < 3; ++a) x[a] = fmin(fmax(x[a], lo[a]), hi[a]);        // clip_guess, main.py:250-252
    double g[3], A[6];
    double cost = solve_eval(mic, pairs, td, weights, P, sp.c, x, g, A, true);
    double lam = 1e-3;
    int it = 0, status = 0;       // status 1: gtol, 2: xtol, 3: ftol
    for (; it < sp.max_iter && !status; ++it) {
      // active set: a coordinate sitting on a bound with the gradient pushing outwards is held fixed; the step is
      // solved in the remaining coordinates (a clipped full step is not the minimiser of the model on the box)
      double gn = 0.0;
      bool act[3];
      for (int a = 0; a < 3; ++a) {
        act[a] = (x[a] <= lo[a] && g[a] > 0.0) || (x[a] >= hi[a] && g[a] < 0.0);
        if (!act[a]) gn = fmax(gn, fabs(g[a]));
      }
      if (gn <= sp.gtol) { status = 1; break; }
      double Ar[6], gr[3];
      Ar[0] = act[0] ? 1.0 : A[0];
      Ar[3] = act[1] ? 1.0 : A[3];
      Ar[5] = act[2] ? 1.0 : A[5];
      Ar[1] = (act[0] || act[1]) ? 0.0 : A[1];
      Ar[2] = (act[0] || act[2]) ? 0.0 : A[2];
      Ar[4] = (act[1] || act[2]) ? 0.0 : A[4];
      for (int a = 0; a < 3; ++a) gr[a] = act[a] ? 0.0 : g[a];
      bool moved = false;
      for (int tries = 0; tries < 40 && !moved && !status; ++tries) {
        double d[3];
        if (!solve3(Ar, gr, lam, d)) { lam = lam * 10.0 + 1e-12; continue; }
        double xn[3];
        for (int a = 0; a < 3; ++a) xn[a] = fmin(fmax(x[a] + d[a], lo[a]), hi[a]);
        double gd[3], Ad[6];
        const double cn = solve_eval(mic, pairs, td, weights, P, sp.c, xn, gd, Ad, false);
        const double step = sqrt((xn[0] - x[0]) * (xn[0] - x[0]) + (xn[1] - x[1]) * (xn[1] - x[1]) + (xn[2] - x[2]) * (xn[2] - x[2]));
        const double xnorm = sqrt(x[0] * x[0] + x[1] * x[1] + x[2] * x[2]);
        if (cn < cost) {
          const double dec = cost - cn;
          for (int a = 0; a < 3; ++a) x[a] = xn[a];
          if (dec <= sp.ftol * cost) status = 3;
          if (step <= sp.xtol * (sp.xtol + xnorm)) status = 2;
          cost = solve_eval(mic, pairs, td, weights, P, sp.c, x, g, A, true);
          lam = fmax(lam / 3.0, 1e-12);
          moved = true;
        } else {
          if (step <= sp.xtol * (sp.xtol + xnorm)) status = 2;       // the model cannot improve on x at this scale
          lam *= 4.0;
        }
      }
      if (!moved && !status) status = 2;
    }
    if (lane == 0) {
      out_pos[3 * s] = x[0]; out_pos[3 * s + 1] = x[1]; out_pos[3 * s + 2] = x[2];
      if (out_cost) out_cost[s] = cost;
      if (out_iter) out_iter[s] = status ? it : -it;
    }
  }
}

}  // namespace pal
