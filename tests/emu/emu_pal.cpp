// Host emulation harness (TEST ONLY): runs the kernel bodies of pyaudiolocalization_b200/csrc
// on OS threads via pal_simt.h's -DPAL_EMU mode so the CPU test-suite can check kernel logic
// against the oracle without a GPU.  Built by tests/emu/build_emu.py into tests/emu/_build/.
#include "pal_pfa4095.cuh"

using namespace pal;

extern "C" {

void emu_fwd4095(const float* sig, int M, long long B, float* spec /* [B][M][2080][2] */, float* hq /* [B][M][2] */, int grid) {
  const long long units = B * ((M + 1) / 2);
  simt::launch(grid, 128, sizeof(FwdSmem), [&](char* smem) {
    fwd4095_body<128>(sig, M, units, reinterpret_cast<cpxf*>(spec), hq, smem);
  });
}

void emu_pair4095_fast(const float* spec, const float* hq, const int* pairs, int M, int P, long long B, int win_half, int dist,
                       float eps, int* k_idx, float* peak, float* gmax, unsigned* flags, float* corr_out,
                       int grid, int phase_sync) {
  constexpr int W = 2;
  const cpxf* sp = reinterpret_cast<const cpxf*>(spec);
  simt::launch(grid, 32 * W, W * sizeof(FastWarpSmem) + 16 + sizeof(int) * 2 * size_t(P), [&](char* smem) {
    if (phase_sync >= 2) {     // TMEM-assisted variant; 2: pair table copied to shared memory, 3: read from global memory
      const int ps = phase_sync == 2 ? 1 : 0;
      if (corr_out)
        pair4095_tmem_body<W, true>(sp, hq, pairs, M, P, B * P, win_half, dist, eps, k_idx, peak, gmax, flags, corr_out, smem, ps);
      else
        pair4095_tmem_body<W, false>(sp, hq, pairs, M, P, B * P, win_half, dist, eps, k_idx, peak, gmax, flags, corr_out, smem, ps);
      return;
    }
    if (corr_out)
      pair4095_fast_body<W, true>(sp, hq, pairs, M, P, B * P, win_half, dist, eps, k_idx, peak, gmax, flags, corr_out, smem);
    else
      pair4095_fast_body<W, false>(sp, hq, pairs, M, P, B * P, win_half, dist, eps, k_idx, peak, gmax, flags, corr_out, smem);
  });
}

// mode: 0 = float from spectra, 1 = float from signals, 2 = double from signals
void emu_pair4095_exact(int mode, const float* sig, const float* spec, const int* pairs, int M, int P, long long B,
                        const int* item_list, const int* item_count, int win_half, int dist, int method,
                        float mult, int num_peaks, int* k_idx, int* k_count, float* peak, float* gmax,
                        unsigned* flags, float* corr_out, int grid) {
  PickParams pp{win_half, dist, method, mult, num_peaks};
  constexpr int NT = 64;
  const cpxf* sp = reinterpret_cast<const cpxf*>(spec);
  if (mode == 0)
    simt::launch(grid, NT, sizeof(ExactSmem<float>), [&](char* smem) {
      pair4095_exact_body<float, NT, true>(sig, sp, pairs, M, P, B * P, item_list, item_count, pp, k_idx,
                                            k_count, peak, gmax, flags, 0u, 0u, corr_out, smem);
    });
  else if (mode == 1)
    simt::launch(grid, NT, sizeof(ExactSmem<float>), [&](char* smem) {
      pair4095_exact_body<float, NT, false>(sig, sp, pairs, M, P, B * P, item_list, item_count, pp, k_idx,
                                             k_count, peak, gmax, flags, 0u, 0u, corr_out, smem);
    });
  else
    simt::launch(grid, NT, sizeof(ExactSmem<double>), [&](char* smem) {
      pair4095_exact_body<double, NT, false>(sig, sp, pairs, M, P, B * P, item_list, item_count, pp, k_idx,
                                              k_count, peak, gmax, flags, PAL_FLAG_REFINED, 0u, corr_out, smem);
    });
}

// stand-alone row pick (any n) for fuzzing the peak selection against the oracle
void emu_peakpick_f64(const double* c, int n, int c0, int win_half, int dist, int method, double mult,
                      int num_peaks, int* out_k, int* out_count, unsigned* out_flags) {
  constexpr int NT = 64;
  struct Sm { PickScratch ps; int ok[16]; double g, p; };
  std::vector<unsigned char> pk(n);
  simt::launch(1, NT, sizeof(Sm), [&](char* smem) {
    Sm* sm = reinterpret_cast<Sm*>(smem);
    PickResult r = peakpick_row<double, NT>(c, n, c0, win_half, dist, method, mult, num_peaks, pk.data(),
                                            &sm->ps, sm->ok, &sm->g, &sm->p);
    simt::sync_block();
    if (simt::tid() == 0) {
      for (int t = 0; t < r.count; ++t) out_k[t] = sm->ok[t];
      *out_count = r.count;
      *out_flags = r.flags;
    }
  });
}
}

// ---------------------------------------------------------------- arbitrary-length (Bluestein) path
#include "pal_bluestein.cuh"
#include <algorithm>

template <typename T>
static void emu_generic_impl(const float* sig, long long B, int Mics, int ld, int n1, int n2, const int* pairs, int P,
                             int win_half, int dist, int method, float mult, int num_peaks, float eps, int* k_idx,
                             int* k_count, float* peak, float* gmax, unsigned* flags, float* corr_out) {
  constexpr int NT = 64;
  constexpr int TC = 4;
  const int n = n1 + n2 - 1;
  const BluePlan p = make_blue_plan(n);
  std::vector<cpx<T>> chirp(n), tw1(p.M1 / 2 + 1), tw2(p.M2 / 2 + 1), twM(p.M), bhat(p.M);
  simt::launch(2, NT, 16, [&](char*) { blue_init_tables_body<T>(p, chirp.data(), tw1.data(), tw2.data(), twM.data()); });
  BlueTables<T> tb{chirp.data(), tw1.data(), tw2.data(), twM.data(), bhat.data()};
  const int tc = std::min(p.M2, TC);
  constexpr int TRW = 4;
  const size_t cs = fft_tile_smem(sizeof(T), p.M1, tc), rs = fft_tile_smem(sizeof(T), p.M2, std::min(p.M1, TRW) + 1);
  simt::launch(2, NT, cs, [&](char* sm) { colpass_fwd_body<T, NT, TC>(p, tb, LoadBhat<T>{p, chirp.data()}, 1, bhat.data(), sm); });
  simt::launch(2, NT, rs, [&](char* sm) { rowpass_body<T, NT, TRW, false, false>(p, tb, 1, bhat.data(), sm); });
  const int CP = (Mics + 1) / 2;
  const long long rows = B * CP, items = B * P, itr = (items + 1) / 2;    // packed: two real sequences per transform
  std::vector<cpx<T>> conv(size_t(std::max(rows, itr)) * p.M), spec(size_t(rows) * n);
  std::vector<T> corr(size_t(items) * n);
  std::vector<float> scales(size_t(B) * Mics * 2);
  simt::launch(2, NT, 64, [&](char* sm) { row_scale_body<NT>(sig, B * Mics, ld, n1, n2, scales.data(), sm); });
  simt::launch(3, NT, cs, [&](char* sm) {
    colpass_fwd_body<T, NT, TC>(p, tb, LoadSignal2<T>{p, chirp.data(), sig, ld, Mics, CP, n1, n2, nullptr, 0, scales.data()}, rows, conv.data(), sm);
  });
  simt::launch(3, NT, rs, [&](char* sm) { rowpass_body<T, NT, TRW, true, false>(p, tb, rows, conv.data(), sm); });
  simt::launch(3, NT, cs, [&](char* sm) {
    colpass_inv_body<T, NT, TC>(p, tb, StoreSpectrum<T>{p, chirp.data(), spec.data()}, rows, conv.data(), sm);
  });
  const LoadPhat2<T> lp{p, chirp.data(), spec.data(), pairs, Mics, CP, P, 0, items, false, scales.data(), 0, nullptr, 0};
  simt::launch(3, NT, cs, [&](char* sm) { colpass_fwd_body<T, NT, TC>(p, tb, lp, itr, conv.data(), sm); });
  simt::launch(3, NT, rs, [&](char* sm) { rowpass_body<T, NT, TRW, true, true>(p, tb, itr, conv.data(), sm); });
  simt::launch(3, NT, cs, [&](char* sm) {
    colpass_inv_body<T, NT, TC>(p, tb, StoreCorr2<T>{p, chirp.data(), corr.data(), items, lp}, itr, conv.data(), sm);
  });
  const int grid = 2;
  std::vector<unsigned char> pk(size_t(grid) * ((n + 15) / 16 * 16));
  simt::launch(grid, NT, sizeof(RowPickSmem), [&](char* sm) {
    pick_rows_body<T, NT>(corr.data(), n, n2 - 1, items, nullptr, win_half, dist, method, mult, num_peaks, eps,
                          pk.data(), k_idx, k_count, peak, gmax, flags, 0u, 0u, corr_out, sm);
  });
}

extern "C" void emu_generic_gcc_phat(int use_double, const float* sig, long long B, int Mics, int ld, int n1, int n2,
                                     const int* pairs, int P, int win_half, int dist, int method, float mult,
                                     int num_peaks, float eps, int* k_idx, int* k_count, float* peak, float* gmax,
                                     unsigned* flags, float* corr_out) {
  if (use_double)
    emu_generic_impl<double>(sig, B, Mics, ld, n1, n2, pairs, P, win_half, dist, method, mult, num_peaks, eps, k_idx,
                             k_count, peak, gmax, flags, corr_out);
  else
    emu_generic_impl<float>(sig, B, Mics, ld, n1, n2, pairs, P, win_half, dist, method, mult, num_peaks, eps, k_idx,
                            k_count, peak, gmax, flags, corr_out);
}

// ---------------------------------------------------------------- second-generation float32 convolution engine
#include "pal_fft2.cuh"
#include "pal_winpick.cuh"

// the arbitrary-length GCC-PHAT path on the compile-time-planned engine (pal_fft2.cuh), plan `plan_id`
// (fft2::plan_dims; its M must be >= 2n - 1); same loaders / storers / pick kernel as emu_generic_impl<float>
extern "C" int emu_fft2_gcc_phat(int plan_id_and_mode, const float* sig, long long B, int Mics, int ld, int n1, int n2,
                                 const int* pairs, int P, int win_half, int dist, int method, float mult,
                                 int num_peaks, float eps, int* k_idx, int* k_count, float* peak, float* gmax,
                                 unsigned* flags, float* corr_out) {
  using T = float;
  constexpr int NT = 64;
  const int n = n1 + n2 - 1;
  // plan_id_and_mode: plan index (or -1: the product's choice); + 100: reduced pick (StoreWin2 + win_pick_rows_body)
  // + 200: the single-CTA kernel (whole convolution in shared memory; plans of at most 16384 points), full rows
  const bool smem_conv = plan_id_and_mode >= 150;
  const bool fast = !smem_conv && plan_id_and_mode >= 50;
  int plan_id = smem_conv ? plan_id_and_mode - 200 : (fast ? plan_id_and_mode - 100 : plan_id_and_mode);
  if (plan_id < 0) plan_id = fft2::choose_plan(n);
  if (plan_id < 0 || plan_id >= fft2::kNumPlans) return -1;
  const fft2::PlanDims pd = fft2::plan_dims(plan_id);
  const long long M = (long long)pd.M1 * pd.M2;
  if (M < 2LL * n - 1) return -2;
  BluePlan p{n, int(M), pd.M1, pd.M2, 0, 0};
  std::vector<cpxf> chirp(n), tw1(pd.M1), tw2(pd.M2), twf(M), bhat(M), scratch(M);
  simt::launch(2, NT, 16, [&](char*) { fft2::init_tables_body(n, pd.M1, pd.M2, chirp.data(), tw1.data(), tw2.data(), twf.data()); });
  fft2::Tables tb{chirp.data(), tw1.data(), tw2.data(), twf.data(), bhat.data()};
  const int CP = (Mics + 1) / 2;
  const long long rows = B * CP, items = B * P, itr = (items + 1) / 2;
  std::vector<cpxf> conv(size_t(std::max(rows, itr)) * M), spec(size_t(rows) * n);
  std::vector<T> corr(size_t(items) * n);
  std::vector<float> scales(size_t(B) * Mics * 2);
  simt::launch(2, NT, 64, [&](char* sm) { row_scale_body<NT>(sig, B * Mics, ld, n1, n2, scales.data(), sm); });
  const LoadSignal2<T> ls{p, chirp.data(), sig, ld, Mics, CP, n1, n2, nullptr, 0, scales.data()};
  const LoadPhat2<T> lp{p, chirp.data(), spec.data(), pairs, Mics, CP, P, 0, items, false, scales.data(), 0, nullptr, 0};
  bool smem_done = false;
  if (smem_conv) {
    fft2::with_plan(plan_id, [&](auto pl) {
      using PL = decltype(pl);
      if constexpr (fft2::SmemConv<PL>::fits) {
        constexpr int NS = 128;            // threads: a multiple of the lane count of both passes
        constexpr size_t sb = fft2::SmemConv<PL>::smem;
        std::vector<cpxf> bs(M);
        StoreRaw<T> none{scratch.data(), M};
        simt::launch(1, NS, sb, [&](char* sm) {
          fft2::conv_smem_body<PL, NS, 2>(tb, nullptr, bs.data(), LoadBhat<T>{p, chirp.data()}, none, 1, sm);
        });
        simt::launch(3, NS, sb, [&](char* sm) {
          fft2::conv_smem_body<PL, NS, 0>(tb, bs.data(), nullptr, ls, StoreSpectrum<T>{p, chirp.data(), spec.data()}, rows, sm);
        });
        simt::launch(3, NS, sb, [&](char* sm) {
          fft2::conv_smem_body<PL, NS, 1>(tb, bs.data(), nullptr, lp, StoreCorr2<T>{p, chirp.data(), corr.data(), items, lp}, itr, sm);
        });
        smem_done = true;
      }
    });
    if (!smem_done) return -3;
  }
  if (!smem_done) fft2::with_plan(plan_id, [&](auto pl) {
    using PL = decltype(pl);
    simt::launch(2, NT, PL::col_smem, [&](char* sm) { fft2::colpass_fwd_body<PL, NT>(tb, LoadBhat<T>{p, chirp.data()}, 1, scratch.data(), sm); });
    simt::launch(2, NT, PL::row_smem, [&](char* sm) { fft2::rowpass_body<PL, NT, 2>(tb, 1, scratch.data(), bhat.data(), sm); });
    simt::launch(3, NT, PL::col_smem, [&](char* sm) { fft2::colpass_fwd_body<PL, NT>(tb, ls, rows, conv.data(), sm); });
    simt::launch(3, NT, PL::row_smem, [&](char* sm) { fft2::rowpass_body<PL, NT, 0>(tb, rows, conv.data(), nullptr, sm); });
    simt::launch(3, NT, PL::col_smem, [&](char* sm) {
      fft2::colpass_inv_body<PL, NT>(tb, StoreSpectrum<T>{p, chirp.data(), spec.data()}, rows, conv.data(), sm);
    });
    if (fast) {
      // the product's fast path: channels unpacked + whitened once, pairs from the half spectra, window pick
      const int Hn = n / 2 + 1;
      std::vector<cpxf> U(size_t(B) * Mics * Hn);
      std::vector<float> hq(size_t(B) * Mics * 2);
      simt::launch(2, NT, 4 * (NT / 32) * sizeof(float), [&](char* sm) {
        whiten_unpack_body<NT>(spec.data(), n, rows, Mics, CP, scales.data(), 0, 0, U.data(), hq.data(), sm);
      });
      const LoadPhatU lu{p, chirp.data(), U.data(), pairs, Mics, P, Hn, 0, items, scales.data(), 0};
      simt::launch(3, NT, PL::col_smem, [&](char* sm) { fft2::colpass_fwd_body<PL, NT>(tb, lu, itr, conv.data(), sm); });
      simt::launch(3, NT, PL::row_smem, [&](char* sm) { fft2::rowpass_body<PL, NT, 1>(tb, itr, conv.data(), nullptr, sm); });
      const WinGeom wg = make_win_geom(n, n2 - 1, win_half, dist, eps);
      const int tiles = PL::M2 / PL::TC;
      std::vector<float> win(size_t(2 * itr) * wg.wstride, -7.f), pmax(size_t(2 * itr) * tiles, -7.f);
      simt::launch(3, NT, PL::col_smem, [&](char* sm) {
        fft2::colpass_inv_body<PL, NT>(tb, StoreWinU{p, chirp.data(), win.data(), pmax.data(), items, lu, wg, tiles}, itr, conv.data(), sm);
      });
      simt::launch(2, NT, 16, [&](char*) {
        win_pick_rows_body<NT>(win.data(), pmax.data(), tiles, items, wg, 0, k_idx, k_count, peak, gmax, flags, 0u,
                               WhitenRef{hq.data(), pairs, Mics, P, 0});
      });
      return;
    }
    simt::launch(3, NT, PL::col_smem, [&](char* sm) { fft2::colpass_fwd_body<PL, NT>(tb, lp, itr, conv.data(), sm); });
    simt::launch(3, NT, PL::row_smem, [&](char* sm) { fft2::rowpass_body<PL, NT, 1>(tb, itr, conv.data(), nullptr, sm); });
    simt::launch(3, NT, PL::col_smem, [&](char* sm) {
      fft2::colpass_inv_body<PL, NT>(tb, StoreCorr2<T>{p, chirp.data(), corr.data(), items, lp}, itr, conv.data(), sm);
    });
  });
  if (fast) return plan_id;
  const int grid = 2;
  std::vector<unsigned char> pk(size_t(grid) * ((n + 15) / 16 * 16));
  simt::launch(grid, NT, sizeof(RowPickSmem), [&](char* sm) {
    pick_rows_body<T, NT>(corr.data(), n, n2 - 1, items, nullptr, win_half, dist, method, mult, num_peaks, eps,
                          pk.data(), k_idx, k_count, peak, gmax, flags, 0u, 0u, corr_out, sm);
  });
  return plan_id;
}

// ---------------------------------------------------------------- stage 1: image sources + renderer
#include "pal_render.cuh"

extern "C" void emu_image_sources(const double* sources, long long n_scenes, const double* planes, const int* plane_mat,
                                  int n_planes, const double* mat_abs, const double* mat_freq, const double* mics,
                                  int n_mics, int max_order, double frequency, double threshold, double round_scale,
                                  int k_max, double* out_pos, int* out_mat, int* out_count) {
  constexpr int NT = 64;
  ImgParams ip{n_planes, n_mics, max_order, k_max, frequency, threshold, round_scale};
  const int cmax = k_max * n_planes;
  const size_t per_block = size_t(cmax) * 3 * 8 * 2 + size_t(k_max + 1) * 3 * 8 + size_t(cmax) * 4 + 64;
  const int grid = 2;
  std::vector<char> scratch(per_block * grid);
  simt::launch(grid, NT, 64 * sizeof(int), [&](char* sm) {
    image_sources_body<NT>(ip, sources, n_scenes, planes, 0, plane_mat, mat_abs, mat_freq, mics, 0, out_pos, out_mat,
                           out_count, scratch.data(), per_block, sm);
  });
}

extern "C" void emu_path_table(const double* src, const double* img_pos, const int* img_mat, int n_img, const double* mics,
                               int n_mics, const double* mat_abs, const double* mat_freq, int air_mat, double frequency,
                               double c_sound, double* tau, double* gain) {
  const int total = n_mics * (n_img + 1);
  simt::launch((total + 63) / 64, 64, 16, [&](char*) {
    path_table_body(src, img_pos, img_mat, n_img, mics, n_mics, mat_abs, mat_freq, air_mat, frequency, c_sound, tau, gain);
  });
}

extern "C" void emu_render_scene(const float* base, int n_base, int N, const double* tau, const double* gain, int n_mics,
                                 int k1, double fs, int n_keep, float* out) {
  using T = float;
  constexpr int NT = 64, TC = 4, J = 4;
  const BluePlan p = make_blue_plan(2 * N);
  std::vector<cpx<T>> chirp(p.n), tw1(p.M1 / 2 + 1), tw2(p.M2 / 2 + 1), twM(p.M), bhat(p.M);
  simt::launch(2, NT, 16, [&](char*) { blue_init_tables_body<T>(p, chirp.data(), tw1.data(), tw2.data(), twM.data()); });
  BlueTables<T> tb{chirp.data(), tw1.data(), tw2.data(), twM.data(), bhat.data()};
  const int tc = std::min(p.M2, TC);
  constexpr int TRW = 4;
  const size_t cs = fft_tile_smem(sizeof(T), p.M1, tc), rs = fft_tile_smem(sizeof(T), p.M2, std::min(p.M1, TRW) + 1);
  simt::launch(2, NT, cs, [&](char* sm) { colpass_fwd_body<T, NT, TC>(p, tb, LoadBhat<T>{p, chirp.data()}, 1, bhat.data(), sm); });
  simt::launch(2, NT, rs, [&](char* sm) { rowpass_body<T, NT, TRW, false, false>(p, tb, 1, bhat.data(), sm); });
  std::vector<cpx<T>> conv(size_t(n_mics) * p.M), X(p.n);
  std::vector<cpxf> G(size_t(n_mics) * (N + 1));
  std::vector<int> live(n_mics);
  simt::launch(3, NT, cs, [&](char* sm) {
    colpass_fwd_body<T, NT, TC>(p, tb, LoadSignal<T>{p, chirp.data(), base, n_base}, 1, conv.data(), sm);
  });
  simt::launch(3, NT, rs, [&](char* sm) { rowpass_body<T, NT, TRW, true, false>(p, tb, 1, conv.data(), sm); });
  simt::launch(3, NT, cs, [&](char* sm) {
    colpass_inv_body<T, NT, TC>(p, tb, StoreSpectrum<T>{p, chirp.data(), X.data()}, 1, conv.data(), sm);
  });
  const size_t ts = 4 * ((k1 + 3) & ~3) + 16 * size_t(k1) + 16;
  const RenderRows rr{tau, gain, nullptr, nullptr, k1, n_mics};
  simt::launch(3, NT, ts, [&](char* sm) { transfer_body<NT, J>(X.data(), N, rr, 0, n_mics, fs, G.data(), live.data(), sm); });
  simt::launch(3, NT, cs, [&](char* sm) {
    colpass_fwd_body<T, NT, TC>(p, tb, LoadHermitian2<T>{p, chirp.data(), G.data(), N, n_mics}, (n_mics + 1) / 2, conv.data(), sm);
  });
  simt::launch(3, NT, rs, [&](char* sm) { rowpass_body<T, NT, TRW, true, true>(p, tb, (n_mics + 1) / 2, conv.data(), sm); });
  const int fade = int(0.01 * N);
  simt::launch(3, NT, cs, [&](char* sm) {
    colpass_inv_body<T, NT, TC>(p, tb, StoreRender2<T>{p, chirp.data(), out, N, n_keep, fade, rr, 0, n_mics, live.data()}, (n_mics + 1) / 2, conv.data(), sm);
  });
  simt::launch(2, NT, 64, [&](char* sm) { normalise_compress_body<NT>(out, n_mics, n_keep, 0.8f, 1e-8f, true, sm); });
}

extern "C" void emu_path_table_batched(const double* sources, const double* img_pos, const int* img_mat, const int* img_count,
                                       long long n_scenes, int k_max, const double* mics, int n_mics, long long mic_stride,
                                       const double* mat_abs, const double* mat_freq, int air_mat, double frequency,
                                       double c_sound, int k_stride, double* tau, double* gain, int* path_count,
                                       double* max_tau) {
  constexpr int NT = 64;
  simt::launch(2, NT, 64, [&](char* sm) {
    path_table_batched_body<NT>(sources, img_pos, img_mat, img_count, n_scenes, k_max, mics, n_mics, mic_stride, mat_abs,
                                mat_freq, air_mat, frequency, c_sound, k_stride, tau, gain, path_count, max_tau, sm);
  });
}

// ---------------------------------------------------------------- channel filter (filtfilt)
#include "pal_filter.cuh"
extern "C" void emu_filtfilt_f64(const double* x, long long n_rows, int n, const double* b, const double* a, const double* zi,
                                 int ntaps, int padlen, double* y) {
  constexpr int NT = 64;
  FiltParams fp{};
  fp.ntaps = ntaps;
  fp.padlen = padlen;
  for (int i = 0; i < kFiltMaxTaps; ++i) {
    fp.b[i] = i < ntaps ? b[i] : 0.0;
    fp.a[i] = i < ntaps ? a[i] : 0.0;
    fp.zi[i] = i < ntaps - 1 ? zi[i] : 0.0;
  }
  std::vector<double> work(size_t((n_rows + 31) / 32) * (n + 2 * padlen) * 32);
  simt::launch(2, NT, size_t(NT / 32) * 32 * 33 * sizeof(double), [&](char* sm) {
    filtfilt_body<double, NT>(x, n_rows, n, fp, work.data(), y, sm);
  });
}

// ---------------------------------------------------------------- channel alignment front end (pal_sync.cuh)
#include "pal_sync.cuh"
extern "C" void emu_sync_align(const double* sig, long long S, int Mics, int ld, const int* lens, int* ref_idx,
                               int* peak_index, double* absmax, double* win, double* energy) {
  using T = double;
  constexpr int NT = 64, TC = 4, TRW = 4;
  const int n = 2 * ld - 1;
  const BluePlan p = make_blue_plan(n);
  std::vector<cpx<T>> chirp(n), tw1(p.M1 / 2 + 1), tw2(p.M2 / 2 + 1), twM(p.M), bhat(p.M);
  simt::launch(2, NT, 16, [&](char*) { blue_init_tables_body<T>(p, chirp.data(), tw1.data(), tw2.data(), twM.data()); });
  BlueTables<T> tb{chirp.data(), tw1.data(), tw2.data(), twM.data(), bhat.data()};
  const int tc = std::min(p.M2, TC);
  const size_t cs = fft_tile_smem(sizeof(T), p.M1, tc), rs = fft_tile_smem(sizeof(T), p.M2, std::min(p.M1, TRW) + 1);
  simt::launch(2, NT, cs, [&](char* sm) { colpass_fwd_body<T, NT, TC>(p, tb, LoadBhat<T>{p, chirp.data()}, 1, bhat.data(), sm); });
  simt::launch(2, NT, rs, [&](char* sm) { rowpass_body<T, NT, TRW, false, false>(p, tb, 1, bhat.data(), sm); });
  simt::launch(2, NT, sizeof(double) * Mics, [&](char* sm) {
    sync_energy_body<NT>(sig, S, Mics, ld, ld, lens, energy, ref_idx, sm);
  });
  const long long rows = S * Mics;
  std::vector<cpx<T>> conv(size_t(rows) * p.M), spec(size_t(rows) * n);
  std::vector<T> corr(size_t(rows) * n);
  simt::launch(3, NT, cs, [&](char* sm) {
    colpass_fwd_body<T, NT, TC>(p, tb, LoadSignalF64<T>{p, chirp.data(), sig, ld, ld, lens, 0}, rows, conv.data(), sm);
  });
  simt::launch(3, NT, rs, [&](char* sm) { rowpass_body<T, NT, TRW, true, false>(p, tb, rows, conv.data(), sm); });
  simt::launch(3, NT, cs, [&](char* sm) {
    colpass_inv_body<T, NT, TC>(p, tb, StoreSpectrum<T>{p, chirp.data(), spec.data()}, rows, conv.data(), sm);
  });
  simt::launch(3, NT, cs, [&](char* sm) {
    colpass_fwd_body<T, NT, TC>(p, tb, LoadCross<T>{p, chirp.data(), spec.data(), ref_idx, 0, Mics, 0}, rows, conv.data(), sm);
  });
  simt::launch(3, NT, rs, [&](char* sm) { rowpass_body<T, NT, TRW, true, true>(p, tb, rows, conv.data(), sm); });
  simt::launch(3, NT, cs, [&](char* sm) {
    colpass_inv_body<T, NT, TC>(p, tb, StoreCorr<T>{p, chirp.data(), corr.data()}, rows, conv.data(), sm);
  });
  simt::launch(2, NT, sizeof(SyncPickSmem), [&](char* sm) {
    sync_pick_body<T, NT>(corr.data(), n, rows, 0, Mics, ref_idx, ld, lens, peak_index, absmax, win, sm);
  });
}

extern "C" void emu_pad_rows_f64(const double* in, long long n_rows, long long ld_in, const int* lens, const int* pad,
                                 double* out, long long ld_out) {
  simt::launch(3, 64, 16, [&](char*) { pad_rows_body<double>(in, n_rows, ld_in, int(ld_in), lens, pad, out, ld_out); });
}

// ---------------------------------------------------------------- batched position solve (pal_solver.cuh)
#include "pal_solver.cuh"
extern "C" void emu_solve_positions(const double* mics, long long mic_stride, int n_mics, const int* pairs, int n_pairs,
                                    const double* tdoa, const double* weights, const double* x0, const double* lo, const double* hi,
                                    long long n_scenes, double c, double buffer, int max_iter, double xtol, double ftol, double gtol,
                                    double* out_pos, double* out_cost, int* out_iter) {
  constexpr int NT = 64;
  const int grid = 2;
  std::vector<double> scratch(size_t(grid) * (NT / 32) * n_pairs);
  const SolveParams sp{n_mics, n_pairs, max_iter, c, buffer, xtol, ftol, gtol};
  simt::launch(grid, NT, 16, [&](char*) {
    solve_positions_body<NT>(sp, mics, mic_stride, pairs, tdoa, weights, x0, lo, hi, n_scenes, scratch.data(), out_pos, out_cost, out_iter);
  });
}
