/* pal_b200.h -- C ABI of libpal_b200.so, the B200-native (sm_100a) implementation of the
 * PyAudioLocalization data-parallel hot path.
 *
 * The reference (zeynelacikgoez/PyAudioLocalization) is pure Python and has no FFI of its own
 * (SURVEY.md section 8b); its interface for this path is a handful of module-level functions.
 * Each entry point below names the reference function(s) it replaces (file:line relative to
 * the reference repository).  The Python host layer (pyaudiolocalization_b200/) binds these
 * with ctypes and mirrors the reference signatures; INTEGRATION.md shows the stub a
 * maintainer of the reference would add.
 *
 * Conventions
 *   - every pointer named *_dev is DEVICE memory owned by the caller (e.g. a torch tensor);
 *     the library never allocates, frees or synchronises; all work is enqueued on `stream`
 *     (a cudaStream_t passed as void*; NULL = legacy default stream)
 *   - return value 0 = success, negative = error; pal_last_error() gives the text
 *     (thread-local, never NULL).  No C++ exception crosses this boundary.
 *   - integer decisions that the reference takes in float64 on the host (window half-width,
 *     peak distance: utils.py:151,163) are taken by the caller and passed as integers.
 */
#ifndef PAL_B200_H
#define PAL_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PAL_ABI_VERSION 8

/* error codes */
#define PAL_OK 0
#define PAL_ERR_INVALID (-1)    /* bad argument (NULL pointer, negative size, unsupported option) */
#define PAL_ERR_WORKSPACE (-2)  /* workspace too small; see pal_gcc_phat_workspace */
#define PAL_ERR_CUDA (-3)       /* a CUDA runtime call or kernel launch failed */
#define PAL_ERR_UNSUPPORTED (-4)

/* per-row flag bits written to flags_dev */
#define PAL_FLAG_NEAR_TIE 1u
#define PAL_FLAG_CHAIN 2u
#define PAL_FLAG_PLATEAU 4u
#define PAL_FLAG_REFINED 8u
#define PAL_FLAG_FALLBACK_ARGMAX 16u
#define PAL_FLAG_ALT_THRESHOLD 32u
#define PAL_FLAG_STACK_OVERFLOW 64u

/* Options of utils.get_time_delays_phat (utils.py:121-127) in integer form. */
typedef struct pal_tdoa_params {
  int32_t win_half;    /* largest |lag| in samples with |lag|/fs <= max_expected_delay (utils.py:163);
                          -1 = max_expected_delay is None, -2 = nothing passes */
  int32_t peak_dist;   /* int(fs*0.001) >= 1 (utils.py:151) */
  int32_t thr_method;  /* 0 = 'median' (and any unknown string, utils.py:148-149), 1 = 'adaptive' */
  float thr_mult;      /* threshold_multiplier */
  int32_t num_peaks;   /* 1..16 */
  float tie_eps;       /* fp32 fast path: decisions closer than this are re-evaluated in float64 */
  int32_t refine;      /* 1 = run the float64 re-evaluation of flagged rows (default), 0 = skip */
  int32_t len_first;   /* valid samples of the first signal of a pair (n1); 0 = n_samples */
  int32_t len_second;  /* valid samples of the second signal (n2); 0 = n_samples.  n1 != n2 is only
                          meaningful for M == 2 (row 0 = sig1, row 1 = sig2, zero beyond their lengths) */
} pal_tdoa_params;

int pal_abi_version(void);
const char* pal_last_error(void);

/* number of kernels this library has launched since load (for bench accounting) */
unsigned long long pal_launch_count(void);

/* Measurement hook: while set (stage != 0) the library records `start_event` right before and
 * `stop_event` right after every launch of the named stage on the stream of that launch, so a
 * caller can time ONE kernel inside the pipeline with CUDA events.  Both are cudaEvent_t
 * created by the caller (timing enabled).  stage: 0 = off, 1 = forward transforms,
 * 2 = fused pair kernel (cross-spectrum of the whitened spectra + inverse DFT + peak pick), 3 = float64
 * re-evaluation of flagged rows.  Process-global; not for concurrent use. */
int pal_profile_hook(int32_t stage, void* start_event, void* stop_event);

/* Bytes of device workspace pal_gcc_phat_tdoa needs to process all B frames in one pass
 * (it accepts less and then walks the batch in chunks; `min_bytes`, if not NULL, receives the
 * smallest usable size). */
int pal_gcc_phat_workspace(int64_t B, int32_t M, int32_t n_samples, int32_t P, size_t* bytes,
                           size_t* min_bytes);

/* Batched GCC-PHAT + TDOA pick.
 * Replaces, for every frame b and mic pair p = (i, j):
 *     utils.phat_correlation(sig[b][i], sig[b][j])          utils.py:108-119
 *     utils.get_time_delays_phat(..., num_peaks, ...)       utils.py:121-181
 *     np.max(corr)                                          main.py:223
 * i.e. the double loop main.py:202-228.
 *   sig_dev   [B][M][n_samples] float32
 *   pairs_dev [P][2] int32 (i, j)
 *   k_idx_dev [B][P][num_peaks] int32: raw IFFT index k of each selected peak, -1 padded;
 *             the reference's time delay is (k - (n_samples-1)) / fs  (utils.py:141-142)
 *   k_count_dev [B][P] int32 or NULL; peak_dev/gmax_dev [B][P] float32: corr[k0], max(corr)
 *   flags_dev [B][P] uint32; corr_opt_dev NULL or [B][P][n1+n2-1] float32 (FFT order)
 * n_samples == 2048 (n = 4095 = 5*7*9*13) takes the fused prime-factor kernels; any other length takes the
 * chirp-z (Bluestein) path.  Neither synchronises: flagged rows are counted and compacted on the device and the
 * float64 sweep over them reads its count there (one exception: a workspace so small that the sweep would need more
 * than 96 rounds makes the Bluestein path read the count back once).
 * Exactness: the float32 kernels flag every row whose decision is closer to an alternative than the row's margin:
 * tie_eps plus a per-row bound on what float32 costs THAT row -- the rounding noise of the forward transforms seen
 * through PHAT's unit weights (large for tonal or band-limited frames whose stop band lies below the float32 noise of
 * the pass band) and the absolute 1e-10 of utils.py:117 (frames below about -69 dBFS).  Rows whose bound threatens the
 * VALUES (more than 1e-4 of max(corr)) are flagged whole.  With refine != 0 flagged rows are re-evaluated in float64
 * from the raw samples, so that lag indices equal the reference's and peak / gmax stay within 1e-4.  num_peaks > 1 is
 * evaluated in float64 throughout.  With refine == 0 flagged rows keep the float32 answer (the chirp-z path then skips
 * the audit altogether and leaves the flags 0): nothing is guaranteed for rows a float64 look would decide differently.
 * INGEST CONTRACT: the rows are float32.  "Equal to the reference" therefore means: equal to the reference run on these
 * float32 samples (up-cast exactly to float64) -- what a capture chain or a float32 renderer delivers.  A caller that
 * holds genuine float64 signals and needs the reference's decision on THEM uses pal_gcc_phat_tdoa_f64 below.
 */
int pal_gcc_phat_tdoa(const float* sig_dev, int64_t B, int32_t M, int32_t n_samples,
                      const int32_t* pairs_dev, int32_t P, const pal_tdoa_params* prm,
                      int32_t* k_idx_dev, int32_t* k_count_dev, float* peak_dev, float* gmax_dev,
                      uint32_t* flags_dev, float* corr_opt_dev, void* ws_dev, size_t ws_bytes,
                      void* stream);

/* 16-bit PCM samples (what capture hardware and WAV files deliver) to float32: out[i] = in[i] * scale.  With
 * scale = 1/32768 this is the float conversion soundfile applies in utils.load_audio_file (utils.py:469), exact in
 * float32.  Lets a caller ship int16 frames over PCIe (half the bytes of float32) and feed pal_gcc_phat_tdoa from
 * out_dev.  in_dev / out_dev: `count` elements, in_dev 2-byte and out_dev 4-byte aligned. */
int pal_pcm16_to_f32(const int16_t* in_dev, int64_t count, float scale, float* out_dev, void* stream);

/* time_lags[k] of utils.py:141-142 for every selected peak: out[i] = (double)(k_idx[i] - (n_second - 1)) / fs
 * with an IEEE round-to-nearest float64 division, i.e. bit-identical to numpy's int64 / float64;
 * NaN where k_idx[i] < 0 (padding).  k_idx_dev / out_dev: `count` elements. */
int pal_tdoa_seconds(const int32_t* k_idx_dev, int64_t count, int32_t n_second, double fs, double* out_dev,
                     void* stream);

/* ---------------------------------------------------------------- stage 1: scene synthesis
 *
 * pal_image_sources replaces utils.generate_image_sources_iterative (utils.py:67-106, with
 * reflect_point_across_plane :29-42, distance :44-48, calculate_attenuation :50-65) for n_scenes
 * independent scenes: float64, the reference's evaluation order, discovery order, the
 * 10^-round_decimals de-duplication key and the mean/min pruning rule.
 *   sources_dev [n_scenes][3] f64; planes_dev [n_planes][4] f64 (a, b, c, d), shared by all scenes
 *   (plane_stride = 0) or per scene (plane_stride = 4*n_planes doubles); plane_mat_dev
 *   [n_planes] index into mat_abs_dev / mat_freq_dev (the 'absorption' / 'freq' table,
 *   materials.py:2-16); mics_dev [n_mics][3] f64 shared by all scenes (mic_stride = 0) or per
 *   scene (mic_stride = 3*n_mics doubles).
 *   out_pos_dev [n_scenes][k_max][3] f64, out_mat_dev [n_scenes][k_max] (material index of the
 *   LAST reflecting plane, utils.py:101), out_count_dev [n_scenes] (-1: more than k_max images).
 */
int pal_image_sources_workspace(int32_t n_planes, int32_t k_max, int64_t n_scenes, size_t* bytes);
int pal_image_sources(const double* sources_dev, int64_t n_scenes, const double* planes_dev, int64_t plane_stride,
                      const int32_t* plane_mat_dev, int32_t n_planes, const double* mat_abs_dev,
                      const double* mat_freq_dev, const double* mics_dev, int32_t n_mics,
                      int64_t mic_stride, int32_t max_order, double frequency, double threshold,
                      int32_t round_decimals, int32_t k_max, double* out_pos_dev, int32_t* out_mat_dev,
                      int32_t* out_count_dev, void* ws_dev, size_t ws_bytes, void* stream);

/* Per (mic, path) delay [s] and raw attenuation of one scene: main.py:94-116.  Path 0 is the
 * direct path (material index air_mat, main.py:108), path k the k-th image source.
 *   tau_dev, gain_dev: [n_mics][n_img + 1] f64 */
int pal_path_table(const double* source_dev, const double* img_pos_dev, const int32_t* img_mat_dev,
                   int32_t n_img, const double* mics_dev, int32_t n_mics, const double* mat_abs_dev,
                   const double* mat_freq_dev, int32_t air_mat, double frequency, double c_sound,
                   double* tau_dev, double* gain_dev, void* stream);

/* Render one scene: the loop main.py:104-122 (sum over paths of fractional_delay(base, tau) *
 * gain, signal_processing.py:66-80; trim; normalize_signal :82-86; dynamic_range_compression
 * :88-94) for all n_mics channels.  N = int((duration + max tau) * fs) is decided by the caller
 * (main.py:102); base_dev holds the n_base = int(fs*duration) source samples (zero-padded to N
 * implicitly); n_keep = n_base when trim_to_duration else N.  out_dev [n_mics][n_keep] f32. */
int pal_render_workspace(int32_t N, int32_t n_mics, size_t* bytes, size_t* min_bytes);
#define PAL_RENDER_NORMALISE_COMPRESS 1  /* apply main.py:121-122; without it the call is the bare
                                          sum of fractional_delay() copies (signal_processing.py:66-80) */
int pal_render_scene(const float* base_dev, int32_t n_base, int32_t N, const double* tau_dev,
                     const double* gain_dev, int32_t n_mics, int32_t n_paths, double fs, int32_t n_keep,
                     int32_t flags, float* out_dev, void* ws_dev, size_t ws_bytes, void* stream);

/* Batched forms for many scenes (BASELINE cfg4 / cfg5).  pal_path_table_batched fills, for every scene s,
 * tau/gain [s][mic][k_stride] (k_stride >= k_max + 1; path 0 direct, path k image k-1 of
 * pal_image_sources' output), path_count[s] = images + 1 (0 if the image list overflowed) and
 * max_tau[s] = the largest delay of the scene, from which the caller forms N = int((duration + max_tau) * fs)
 * exactly as main.py:94-102 does.  Scenes that share N are then rendered together by pal_render_scenes:
 * scene_index_dev [n_bucket_scenes] lists them (NULL = all scenes 0..n-1); out_dev is the whole batch
 * [scenes][n_mics][n_keep] and only the listed scenes' rows are written (not normalised: apply
 * pal_normalise_compress to the batch afterwards, main.py:121-122). */
int pal_path_table_batched(const double* sources_dev, const double* img_pos_dev, const int32_t* img_mat_dev,
                           const int32_t* img_count_dev, int64_t n_scenes, int32_t k_max, const double* mics_dev,
                           int32_t n_mics, int64_t mic_stride, const double* mat_abs_dev, const double* mat_freq_dev,
                           int32_t air_mat, double frequency, double c_sound, int32_t k_stride, double* tau_dev,
                           double* gain_dev, int32_t* path_count_dev, double* max_tau_dev, void* stream);
int pal_render_scenes_workspace(int32_t N, int64_t n_rows, size_t* bytes, size_t* min_bytes);
int pal_render_scenes(const float* base_dev, int32_t n_base, int32_t N, const double* tau_dev, const double* gain_dev,
                      const int32_t* path_count_dev, int32_t k_stride, const int64_t* scene_index_dev,
                      int64_t n_bucket_scenes, int32_t n_mics, double fs, int32_t n_keep, float* out_dev, void* ws_dev,
                      size_t ws_bytes, void* stream);

/* Plans.  Everything pal_render_scenes derives from (N, base signal) alone -- chirp and twiddle tables, the chirp
 * spectrum, the spectrum of the zero-padded base signal (signal_processing.py:69) -- can be built once per N with
 * pal_render_plan into caller-owned memory (pal_render_plan_bytes; scratch_dev: pal_render_plan_scratch_bytes, free
 * again when the call's work has run) and reused by pal_render_scenes_planned for every later bucket of that N and
 * that base signal.  With random rooms almost every scene has its own N, and the plan is about half of the GPU work of
 * a small bucket.  Results are bit-identical to pal_render_scenes.  ws_dev of the planned call: pal_render_rows_workspace. */
int pal_render_plan_bytes(int32_t N, size_t* plan_bytes, size_t* scratch_bytes);
int pal_render_plan(const float* base_dev, int32_t n_base, int32_t N, void* plan_dev, size_t plan_bytes, void* scratch_dev,
                    size_t scratch_bytes, void* stream);
int pal_render_rows_workspace(int32_t N, int64_t n_rows, size_t* bytes, size_t* min_bytes);
int pal_render_scenes_planned(const void* plan_dev, size_t plan_bytes, int32_t n_base, int32_t N, const double* tau_dev,
                              const double* gain_dev, const int32_t* path_count_dev, int32_t k_stride,
                              const int64_t* scene_index_dev, int64_t n_bucket_scenes, int32_t n_mics, double fs,
                              int32_t n_keep, float* out_dev, void* ws_dev, size_t ws_bytes, void* stream);

/* In place on n_rows rows of n float32: mode 0 = normalize_signal (signal_processing.py:82-86),
 * mode 1 = dynamic_range_compression(threshold, epsilon) (signal_processing.py:88-94). */
int pal_normalise_compress(float* rows_dev, int64_t n_rows, int32_t n, float threshold, float epsilon,
                           int32_t mode, void* stream);

/* ---------------------------------------------------------------- between the stages: channel filter
 *
 * Zero-phase IIR filtering of n_rows channels: scipy.signal.filtfilt(b, a, x) with its defaults (method
 * "pad", odd extension of padlen = 3 * ntaps samples, lfilter_zi initial conditions), i.e. what
 * signal_processing.noise_reduction(signal, fs, 'butterworth') (signal_processing.py:124-128) applies to
 * every channel in main.py:191.  b, a, zi are HOST arrays of ntaps, ntaps, ntaps-1 doubles
 * (ntaps = max(len(a), len(b)) <= 16, shorter one zero-padded, a[0] == 1, zi = lfilter_zi(b, a)).
 * x_dev / y_dev: [n_rows][n] float64 (io_f32 == 0) or float32 (io_f32 == 1; the arithmetic is float64
 * either way); n > padlen.  In float64 the result is bit-identical to scipy on x86 (same operation order,
 * no FMA contraction).  Workspace: pal_filtfilt_workspace bytes. */
int pal_filtfilt_workspace(int64_t n_rows, int32_t n, int32_t padlen, size_t* bytes);
int pal_filtfilt(const void* x_dev, int64_t n_rows, int32_t n, int32_t io_f32, const double* b, const double* a,
                 const double* zi, int32_t ntaps, int32_t padlen, void* y_dev, void* ws_dev, size_t ws_bytes, void* stream);

/* ---------------------------------------------------------------- between the stages: channel alignment
 *
 * Front end of utils.synchronize_signals_improved (utils.py:407-457) for n_scenes scenes of n_ch channels
 * (rows of `ld` float64 samples, lens_dev[scene][ch] of them valid, NULL = all ld):
 *   ref_idx_dev[scene]        = np.argmax of the channel energies (first maximum)              utils.py:415-416
 *   for every channel: corr   = scipy.signal.correlate(channel, reference, mode='full')        utils.py:426
 *   peak_index_dev[scene][ch] = np.argmax(np.abs(corr)) in 'full' order (first maximum)         utils.py:427
 *   absmax_dev[scene][ch]     = |corr[peak_index]|; the entry of the reference channel itself is ref_peak, :418-419
 *   win_dev[scene][ch][5]     = corr[peak_index-2 .. peak_index+2] (NaN outside the row): the samples of the
 *                               cubic spline of utils.py:431-437, which the caller evaluates in float64
 *   energy_dev[scene][ch]     optional (NULL = not wanted)
 * All arithmetic is float64 (exact length-(2 ld - 1) transforms).  The library never synchronises. */
int pal_sync_align_workspace(int64_t n_scenes, int32_t n_ch, int32_t ld, size_t* bytes, size_t* min_bytes);
int pal_sync_align(const double* sig_dev, int64_t n_scenes, int32_t n_ch, int32_t ld, const int32_t* lens_dev,
                   int32_t* ref_idx_dev, int32_t* peak_index_dev, double* absmax_dev, double* win_dev,
                   double* energy_dev, void* ws_dev, size_t ws_bytes, void* stream);

/* The padding of utils.py:448-457 (np.pad): out[row][pad_left[row] + j] = in[row][j] for j < lens[row]
 * (NULL = ld_in), zero elsewhere; rows of float64 (io_f32 == 0) or float32 (io_f32 == 1). */
int pal_pad_rows(const void* in_dev, int64_t n_rows, int64_t ld_in, const int32_t* lens_dev, const int32_t* pad_left_dev,
                 void* out_dev, int64_t ld_out, int32_t io_f32, void* stream);

/* Many buckets per launch.  With one room per scene nearly every scene has its own transform length (main.py:102), and
 * one pal_render_scenes_planned call per bucket leaves a sweep bound by launch overhead (thousands of 16-block grids).
 * This entry point renders ALL buckets of a batch: bucket i holds the scenes scene_index_dev[first[i] .. first[i+1]) which
 * share N_of_bucket[i] and whose plan (pal_render_plan) lives at plan_dev_of_bucket[i].  The three arrays are HOST arrays
 * (n_buckets, n_buckets, n_buckets + 1 entries); everything else as in pal_render_scenes_planned.  Buckets that share a
 * convolution plan are issued together, four launches per group, as many buckets per group as the workspace holds.
 * PAL_ERR_UNSUPPORTED: some bucket has no compile-time convolution plan (4 N - 1 > 262144 or 2 N < 1025) or does not fit
 * the workspace on its own -- render that batch bucket by bucket.  Output is identical to the per-bucket calls. */
int pal_render_scenes_grouped(int32_t n_buckets, const void* const* plan_dev_of_bucket, const int32_t* N_of_bucket,
                              const int64_t* first_scene_of_bucket, const double* tau_dev, const double* gain_dev,
                              const int32_t* path_count_dev, int32_t k_stride, const int64_t* scene_index_dev, int32_t n_mics,
                              double fs, int32_t n_keep, float* out_dev, void* ws_dev, size_t ws_bytes, void* stream);

/* FLOAT64 INGEST.  pal_gcc_phat_tdoa takes float32 rows: a caller that holds float64 signals (the reference computes
 * in float64 throughout: utils.py:113-118; main.py:191 hands the float64 output of filtfilt to the pair loop) would
 * have to round them first, and a float64 re-evaluation that starts from rounded samples cannot restore the reference's
 * decision at a near tie -- nor its VALUES for band-limited signals, whose stop-band bins lie below the float32
 * quantisation noise and get unit weight from PHAT.  This entry point reads float64 rows [B][M][n_samples] and runs the
 * float64 kernels (complete find_peaks emulation, utils.py:140-181) on EVERY (frame, pair): same outputs and workspace
 * as pal_gcc_phat_tdoa, every row carries PAL_FLAG_REFINED, prm->tie_eps / prm->refine are ignored.  It is the path of
 * the drop-in single-call functions (utils.get_time_delays_phat, utils.phat_correlation, localize_sound_source); the
 * batched float32 entry point is the throughput path and states its contract above: lags are the reference's for
 * float32-representable inputs. */
int pal_gcc_phat_tdoa_f64(const double* sig_dev, int64_t B, int32_t M, int32_t n_samples, const int32_t* pairs_dev, int32_t P,
                          const pal_tdoa_params* prm, int32_t* k_idx_dev, int32_t* k_count_dev, float* peak_dev, float* gmax_dev,
                          uint32_t* flags_dev, float* corr_opt_dev, void* ws_dev, size_t ws_bytes, void* stream);

/* The kernels of this library size their (persistent) grids by the number of SMs, one resident block -- or as many as
 * fit -- per SM.  A collective that must run WHILE they run (the all-gather of the previous step's lag indices, SURVEY.md
 * section 8e) then finds no SM to start on until a whole kernel has drained, which turns every step boundary into a
 * cross-rank synchronisation point.  pal_reserve_sms(n) makes every later call size its grids for (SM count - n) SMs;
 * 0 (the default) gives the kernels the whole device.  Process-wide; no reference counterpart (the reference is
 * single-process). */
int pal_reserve_sms(int32_t n_sms);

/* ---- batched position solve of a scene sweep ---------------------------------------
 * Replaces, for MANY scenes at once, what main.py:246-274 does per scene on the host: the bounded least-squares fit of
 * the source position to the pair TDOAs with the residuals of utils.py:384-405,
 *     r_p = ((|x - m_j| - |x - m_i|) - c * td_p) * w_p,
 * inside the box of utils.py:364-382 (dynamic_bounds_extended: microphone extent +- (buffer + max(percentile75(c |td|), 1))),
 * which is formed on the device when lo_dev / hi_dev are NULL.  One warp per scene, float64, Levenberg-Marquardt with the
 * step clipped to the box; x0_dev NULL starts at the array centroid (clipped like main.py:250-252).
 *   mics_dev [S or 1][n_mics][3] (mic_stride = 0: one array shared by all scenes, else 3 * n_mics), pairs_dev [P][2],
 *   tdoa_dev [S][P] seconds (pal_tdoa_seconds of the lag indices), weights_dev [P] or NULL (utils.py:484-497),
 *   out_pos_dev [S][3], out_cost_dev [S] or NULL (0.5 sum r^2, scipy's `cost`), out_iter_dev [S] or NULL (iterations;
 *   negative: stopped by max_iter).  xtol / ftol / gtol have scipy.optimize.least_squares' meaning.
 * `localize_sound_source` keeps the reference's host solver (clustering starts, scipy trf, Differential Evolution). */
int pal_solve_positions_workspace(int32_t n_pairs, size_t* bytes);
int pal_solve_positions(const double* mics_dev, int64_t mic_stride, int32_t n_mics, const int32_t* pairs_dev, int32_t n_pairs,
                        const double* tdoa_dev, const double* weights_dev, const double* x0_dev, const double* lo_dev,
                        const double* hi_dev, int64_t n_scenes, double c_sound, double buffer, int32_t max_iter, double xtol,
                        double ftol, double gtol, double* out_pos_dev, double* out_cost_dev, int32_t* out_iter_dev, void* ws_dev,
                        size_t ws_bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* PAL_B200_H */
