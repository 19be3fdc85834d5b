// pal_capi.cu -- __global__ wrappers and the extern "C" boundary of libpal_b200.so.
// Build: nvcc -std=c++17 -O3 -lineinfo -gencode arch=compute_100a,code=sm_100a -shared -Xcompiler -fPIC
#include <cuda_runtime.h>

#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

#include "../../include/pal_b200.h"
#include "pal_pfa4095.cuh"
#include "pal_generic_host.cuh"
#include "pal_render_host.cuh"
#include "pal_filter.cuh"
#include "pal_solver.cuh"

using namespace pal;

namespace {

thread_local std::string g_err;
std::atomic<unsigned long long> g_launches{0};
}  // namespace
namespace palhost {
std::atomic<unsigned long long>* g_launch_counter = &g_launches;
}
namespace {
int g_prof_stage = 0;
cudaEvent_t g_prof_start = nullptr, g_prof_stop = nullptr;
struct ProfScope {   // records the caller's events around one stage (pal_profile_hook)
  bool on;
  cudaStream_t s;
  ProfScope(int stage, cudaStream_t st) : on(g_prof_stage == stage && g_prof_start && g_prof_stop), s(st) {
    if (on) cudaEventRecord(g_prof_start, s);
  }
  ~ProfScope() {
    if (on) cudaEventRecord(g_prof_stop, s);
  }
};

int fail(int code, const std::string& msg) {
  g_err = msg;
  return code;
}
int cuda_fail(cudaError_t e, const char* what) {
  return fail(PAL_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
}
#define PAL_CUDA(call)                                  \
  do {                                                  \
    cudaError_t e_ = (call);                            \
    if (e_ != cudaSuccess) return cuda_fail(e_, #call); \
  } while (0)

#ifndef PAL_FWD_THREADS
#define PAL_FWD_THREADS 256
#endif
constexpr int kFwdThreads = PAL_FWD_THREADS;
#ifndef PAL_FAST_WARPS
#define PAL_FAST_WARPS 8
#endif
constexpr int kFastWarps = PAL_FAST_WARPS;   // warps (= mic pairs in flight) per CTA of the fused pair kernel
constexpr int kExactThreads = 256;
__global__ void __launch_bounds__(kFwdThreads) k_fwd4095(const float* __restrict__ sig, int M, long long units,
                                                       cpxf* __restrict__ spec, float* __restrict__ hq) {
  extern __shared__ __align__(128) char smem[];
  fwd4095_body<kFwdThreads>(sig, M, units, spec, hq, smem);
}

template <bool WRITE_CORR>
#ifdef PAL_FAST_MAXNREG   // tuning experiment
__global__ void __launch_bounds__(kFastWarps * 32, 1) __maxnreg__(PAL_FAST_MAXNREG)
#else
__global__ void __launch_bounds__(kFastWarps * 32, 1)
#endif
    k_pair4095_fast(const cpxf* __restrict__ spec, const float* __restrict__ hq, const int* __restrict__ pairs, int M, int P,
                    long long n_items,
                    int win_half, int dist, float eps, int* __restrict__ k_idx, float* __restrict__ peak,
                    float* __restrict__ gmax, unsigned* __restrict__ flags, float* __restrict__ corr_out) {
  extern __shared__ __align__(128) char smem[];
  pair4095_fast_body<kFastWarps, WRITE_CORR>(spec, hq, pairs, M, P, n_items, win_half, dist, eps, k_idx, peak, gmax,
                                             flags, corr_out, smem);
}

#ifndef PAL_TMEM_WARPS
#define PAL_TMEM_WARPS 12
#endif
constexpr int kTmemWarps = PAL_TMEM_WARPS;   // warps per CTA of the TMEM-assisted pair kernel (3 per scheduler)
template <bool WRITE_CORR>
__global__ void __launch_bounds__(kTmemWarps * 32, 1)
    k_pair4095_tmem(const cpxf* __restrict__ spec, const float* __restrict__ hq, const int* __restrict__ pairs, int M, int P,
                    long long n_items,
                    int win_half, int dist, float eps, int* __restrict__ k_idx, float* __restrict__ peak,
                    float* __restrict__ gmax, unsigned* __restrict__ flags, float* __restrict__ corr_out, int pairs_in_smem) {
  extern __shared__ __align__(128) char smem[];
  pair4095_tmem_body<kTmemWarps, WRITE_CORR>(spec, hq, pairs, M, P, n_items, win_half, dist, eps, k_idx, peak, gmax, flags,
                                             corr_out, smem, pairs_in_smem);
}
// Which fused pair kernel runs (read once): PAL_PAIR_KERNEL=tmem (default; 12 warps per SM, register
// tiles parked in tensor memory; measured 10 % faster on B200) or PAL_PAIR_KERNEL=regs (8 warps per SM,
// everything in registers; also used automatically when the tensor-memory kernel cannot be launched).
bool use_tmem_kernel() {
  static const bool on = [] {
    const char* e = std::getenv("PAL_PAIR_KERNEL");
    return e ? (e[0] == 't') : true;
  }();
  return on;
}

template <typename T, bool FROM_SPECTRA, typename TS = float>
__global__ void __launch_bounds__(kExactThreads, 3)
    k_pair4095_exact(const TS* __restrict__ sig, const cpxf* __restrict__ spec, const int* __restrict__ pairs,
                     int M, int P, long long n_items, const int* __restrict__ item_list,
                     const int* __restrict__ item_count, PickParams pp, int* k_idx, int* k_count, float* peak,
                     float* gmax, unsigned* flags, unsigned extra_flag, unsigned keep_mask, float* corr_out) {
  extern __shared__ __align__(128) char smem[];
  pair4095_exact_body<T, kExactThreads, FROM_SPECTRA, TS>(sig, spec, pairs, M, P, n_items, item_list, item_count, pp,
                                                      k_idx, k_count, peak, gmax, flags, extra_flag, keep_mask,
                                                      corr_out, smem);
}

// rows whose fast-path decision was flagged -> compact list for the float64 kernel
__global__ void k_compact_flagged(const unsigned* __restrict__ flags, long long n_items, long long item0,
                                  unsigned mask, int* __restrict__ list, int* __restrict__ count) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n_items && (flags[i] & mask)) list[atomicAdd(count, 1)] = int(item0 + i);
}
__global__ void k_fill_count(int* k_count, long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) k_count[i] = 1;
}

__global__ void k_tdoa_seconds(const int* __restrict__ k_idx, long long n, int c0, double fs, double* __restrict__ out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    const int k = k_idx[i];
    out[i] = (k < 0) ? __longlong_as_double(0x7ff8000000000000LL) : __ddiv_rn(double(k - c0), fs);
  }
}

// 8 samples per thread: one 16-byte load, two 16-byte stores
__global__ void __launch_bounds__(256) k_pcm16_to_f32(const short* __restrict__ in, long long count, float scale,
                                                      float* __restrict__ out, int head) {
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long gsz = (long long)gridDim.x * blockDim.x;
  // `head` samples bring `in` to 16-byte alignment (and `out` to 32-byte alignment relative to its own base)
  for (long long i = gid; i < head && i < count; i += gsz) out[i] = float(in[i]) * scale;
  const long long body = (count - head) > 0 ? (count - head) / 8 : 0;
  const int4* in8 = reinterpret_cast<const int4*>(in + head);
  for (long long v = gid; v < body; v += gsz) {
    const int4 q = in8[v];
    float o[8];
    const int w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      o[2 * k] = float(short(w[k] & 0xffff)) * scale;
      o[2 * k + 1] = float(short(w[k] >> 16)) * scale;
    }
    float* dst = out + head + 8 * v;
    if ((reinterpret_cast<uintptr_t>(dst) & 15u) == 0) {
      reinterpret_cast<float4*>(dst)[0] = make_float4(o[0], o[1], o[2], o[3]);
      reinterpret_cast<float4*>(dst)[1] = make_float4(o[4], o[5], o[6], o[7]);
    } else {
#pragma unroll
      for (int k = 0; k < 8; ++k) dst[k] = o[k];
    }
  }
  for (long long i = head + 8 * body + gid; i < count; i += gsz) out[i] = float(in[i]) * scale;
}

constexpr int kFiltThreads = 128;
template <typename TIO>
__global__ void __launch_bounds__(kFiltThreads) k_filtfilt(const TIO* __restrict__ x, long long n_rows, int n, FiltParams fp,
                                                           double* __restrict__ work, TIO* __restrict__ y) {
  extern __shared__ __align__(16) char smem[];
  filtfilt_body<TIO, kFiltThreads>(x, n_rows, n, fp, work, y, smem);
}

constexpr int kSolveThreads = 128;
__global__ void __launch_bounds__(kSolveThreads) k_solve_positions(SolveParams sp, const double* mics, long long mic_stride,
                                                                  const int* pairs, const double* tdoa, const double* weights,
                                                                  const double* x0, const double* lo, const double* hi,
                                                                  long long n_scenes, double* scratch, double* out_pos,
                                                                  double* out_cost, int* out_iter) {
  solve_positions_body<kSolveThreads>(sp, mics, mic_stride, pairs, tdoa, weights, x0, lo, hi, n_scenes, scratch, out_pos, out_cost,
                                      out_iter);
}
inline int solve_grid(int sms) { return 8 * sms; }

struct DevInfo {
  int sms = 0;
  int dev = -1;
  size_t smem_optin = 0;     // largest dynamic shared-memory size a block may opt in to
};
std::atomic<int> g_reserved_sms{0};      // pal_reserve_sms: SMs the persistent grids leave to other work (a collective)
int device_info(DevInfo& d) {
  PAL_CUDA(cudaGetDevice(&d.dev));
  PAL_CUDA(cudaDeviceGetAttribute(&d.sms, cudaDevAttrMultiProcessorCount, d.dev));
  int optin = 0;
  PAL_CUDA(cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, d.dev));
  d.smem_optin = size_t(optin);
  const int keep = g_reserved_sms.load(std::memory_order_relaxed);
  if (keep > 0 && d.sms > 2 * keep) d.sms -= keep;
  return PAL_OK;
}

size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// tail of the flagged-row list in the workspace: the device-side counter followed by the per-round counts of the
// device-counted float64 sweep (2 x kMaxDevRounds ints)
constexpr size_t kListTail = 1024;
static_assert((1 + 2 * palhost::kMaxDevRounds) * sizeof(int) <= kListTail, "list tail too small");
constexpr unsigned kRefineMask = PAL_FLAG_NEAR_TIE | PAL_FLAG_CHAIN | PAL_FLAG_PLATEAU;

// arbitrary-length path (Bluestein): float32 sweep with a near-tie audit, then a float64 sweep
// over the flagged rows only
int generic_tdoa(const float* sig_dev, int64_t B, int32_t M, int32_t n_samples, int n1, int n2, const int32_t* pairs_dev,
                 int32_t P, const pal_tdoa_params* prm, int32_t* k_idx_dev, int32_t* k_count_dev, float* peak_dev,
                 float* gmax_dev, uint32_t* flags_dev, float* corr_opt_dev, void* ws_dev, size_t ws_bytes,
                 cudaStream_t stream, const double* sig64_dev = nullptr) {
  if (B * (int64_t)P > 0x3fffffffLL) return fail(PAL_ERR_INVALID, "pal_gcc_phat_tdoa: B*P too large; split the batch");
  if (reinterpret_cast<uintptr_t>(ws_dev) & 255u) return fail(PAL_ERR_INVALID, "pal_gcc_phat_tdoa: ws_dev must be 256-byte aligned");
  if (n1 + n2 - 1 < 3) return fail(PAL_ERR_INVALID, "pal_gcc_phat_tdoa: signals too short (n1+n2-1 < 3)");
  DevInfo di;
  if (int rc = device_info(di)) return rc;
  const size_t list_bytes = align_up(size_t(B) * P * sizeof(int), 256) + kListTail;
  const size_t rows_bytes = align_up(size_t(B) * P * 2 * sizeof(int), 256);
  const size_t scale_bytes = 2 * align_up(size_t(B) * M * 2 * sizeof(float), 256);     // scales [B*M][2], then bounds [B*M][2]
  const int n = n1 + n2 - 1;
  if (ws_bytes < list_bytes + rows_bytes + scale_bytes + std::max(palhost::generic_min_bytes<float>(n, M, di.sms),
                                                   palhost::generic_min_bytes<double>(n, 2, di.sms)))
    return fail(PAL_ERR_WORKSPACE, "pal_gcc_phat_tdoa: workspace too small");
  char* ws = static_cast<char*>(ws_dev);
  int* list = reinterpret_cast<int*>(ws);
  int* count = reinterpret_cast<int*>(ws + list_bytes - kListTail);
  int* rows = reinterpret_cast<int*>(ws + list_bytes);
  float* scales = reinterpret_cast<float*>(ws + list_bytes + rows_bytes);
  char* region = ws + list_bytes + rows_bytes + scale_bytes;
  const size_t region_bytes = ws_bytes - list_bytes - rows_bytes - scale_bytes;
  palhost::GenericCall c{sig_dev, (long long)B, M, n_samples, n1, n2, pairs_dev, P,
                         PickParams{prm->win_half, prm->peak_dist, prm->thr_method, prm->thr_mult, prm->num_peaks},
                         prm->refine ? prm->tie_eps : 0.f, k_idx_dev, k_count_dev, peak_dev, gmax_dev, flags_dev,
                         corr_opt_dev, stream, di.sms, scales};
  c.hq = scales + scale_bytes / (2 * sizeof(float));
  // num_peaks > 1 is not the hot path (main.py:204 always asks for one peak): the audit only covers the first peak,
  // so every row is handed to the float64 sweep, as on the n = 4095 path
  if (sig64_dev) {
    // float64 rows: ONE float64 sweep over every item, complete find_peaks emulation -- no float32 stage at all
    c.sig64 = sig64_dev;
    c.eps = 0.f;
    cudaError_t e64 = palhost::run_generic<double>(c, region, region_bytes, nullptr, 0, nullptr, PAL_FLAG_REFINED, 0u);
    if (e64 != cudaSuccess) return cuda_fail(e64, "generic float64 sweep");
    PAL_CUDA(cudaGetLastError());
    return PAL_OK;
  }
  const unsigned all_rows = (prm->refine && prm->num_peaks != 1) ? PAL_FLAG_NEAR_TIE : 0u;
  cudaError_t e = palhost::run_generic<float>(c, region, region_bytes, nullptr, 0, nullptr, all_rows, 0u);
  if (e != cudaSuccess) return cuda_fail(e, "generic float sweep");
  if (!prm->refine) return PAL_OK;
  PAL_CUDA(cudaMemsetAsync(count, 0, sizeof(int), stream));
  const long long n_items = (long long)B * P;
  k_compact_flagged<<<(unsigned)((n_items + 255) / 256), 256, 0, stream>>>(flags_dev, n_items, 0, PAL_FLAG_NEAR_TIE, list, count);
  ++g_launches;
  // float64 re-evaluation of the flagged rows.  The list length stays on the device: the sweep is issued in rounds
  // sized for the worst case, and rounds beyond the real count find nothing to do (a few microseconds each).  Only
  // when the workspace is so small that this would take more than kMaxDevRounds rounds is the count read back.
  {
    palhost::k_rows_of_items<<<(unsigned)((n_items + 255) / 256), 256, 0, stream>>>(list, count, pairs_dev, M, P, rows);
    ++g_launches;
    palhost::GenericCall cd = c;
    cd.eps = 0.f;
    cd.corr_out = nullptr;
    int* dev_items = count + 1;
    int* dev_packed = count + 1 + palhost::kMaxDevRounds;
    e = palhost::run_generic<double>(cd, region, region_bytes, list, (int)n_items, rows, PAL_FLAG_REFINED, kRefineMask, count,
                                     dev_items, dev_packed);
    if (e == cudaSuccess) {
      PAL_CUDA(cudaGetLastError());
      return PAL_OK;
    }
    if (e != cudaErrorNotSupported) return cuda_fail(e, "generic float64 sweep");
    (void)cudaGetLastError();
  }
  int h_count = 0;
  PAL_CUDA(cudaMemcpyAsync(&h_count, count, sizeof(int), cudaMemcpyDeviceToHost, stream));
  PAL_CUDA(cudaStreamSynchronize(stream));
  if (h_count > 0) {
    palhost::k_rows_of_items<<<(h_count + 255) / 256, 256, 0, stream>>>(list, count, pairs_dev, M, P, rows);
    ++g_launches;
    c.eps = 0.f;
    c.corr_out = nullptr;
    e = palhost::run_generic<double>(c, region, region_bytes, list, h_count, rows, PAL_FLAG_REFINED, kRefineMask);
    if (e != cudaSuccess) return cuda_fail(e, "generic float64 sweep");
  }
  PAL_CUDA(cudaGetLastError());
  return PAL_OK;
}

}  // namespace

extern "C" {

int pal_abi_version(void) { return PAL_ABI_VERSION; }
const char* pal_last_error(void) { return g_err.c_str(); }
unsigned long long pal_launch_count(void) { return g_launches.load(); }
int pal_reserve_sms(int32_t n_sms) {
  if (n_sms < 0 || n_sms > 64) return fail(PAL_ERR_INVALID, "pal_reserve_sms: need 0 <= n_sms <= 64");
  g_reserved_sms.store(n_sms);
  return PAL_OK;
}
int pal_profile_hook(int32_t stage, void* start_event, void* stop_event) {
  if (stage < 0 || stage > 3) return fail(PAL_ERR_INVALID, "pal_profile_hook: stage must be 0..3");
  g_prof_stage = stage;
  g_prof_start = static_cast<cudaEvent_t>(start_event);
  g_prof_stop = static_cast<cudaEvent_t>(stop_event);
  return PAL_OK;
}

int pal_gcc_phat_workspace(int64_t B, int32_t M, int32_t n_samples, int32_t P, size_t* bytes, size_t* min_bytes) {
  if (B < 0 || M < 2 || P < 1 || n_samples < 1 || !bytes) return fail(PAL_ERR_INVALID, "pal_gcc_phat_workspace: bad argument");
  const size_t list = align_up(size_t(B) * P * sizeof(int), 256) + kListTail;
  if (n_samples == kFrame2048) {
    const size_t per_frame = align_up(size_t(M) * kSpecSlots * sizeof(cpxf), 256);
    const size_t hq_frame = size_t(M) * 2 * sizeof(float);     // per channel: whitening bound h, rounding-noise term q
    *bytes = (per_frame + hq_frame) * size_t(B > 0 ? B : 1) + list + 256;
    if (min_bytes) *min_bytes = per_frame + hq_frame + list + 256;
    return PAL_OK;
  }
  // Bluestein path: sized for the longest transform the rows allow (n <= 2*n_samples-1)
  int sms = 148;
  DevInfo di;
  if (device_info(di) == PAL_OK && di.sms > 0) sms = di.sms;
  const int n = 2 * n_samples - 1;
  const size_t rows = align_up(size_t(B) * P * 2 * sizeof(int), 256) + 2 * align_up(size_t(B > 0 ? B : 1) * M * 2 * sizeof(float), 256);
  const size_t dmin = palhost::generic_min_bytes<double>(n, 2, sms);
  const size_t fmin = palhost::generic_min_bytes<float>(n, M, sms);
  const size_t ffull = palhost::generic_full_bytes<float>(n, B > 0 ? B : 1, M, P, sms);
  *bytes = list + rows + std::max(ffull, dmin * 4);
  if (min_bytes) *min_bytes = list + rows + std::max(fmin, dmin);
  return PAL_OK;
}

int pal_gcc_phat_tdoa(const float* sig_dev, int64_t B, int32_t M, int32_t n_samples, const int32_t* pairs_dev,
                      int32_t P, const pal_tdoa_params* prm, int32_t* k_idx_dev, int32_t* k_count_dev,
                      float* peak_dev, float* gmax_dev, uint32_t* flags_dev, float* corr_opt_dev, void* ws_dev,
                      size_t ws_bytes, void* stream_) {
  if (!prm) return fail(PAL_ERR_INVALID, "pal_gcc_phat_tdoa: prm is NULL");
  if (B < 0 || M < 2 || P < 1) return fail(PAL_ERR_INVALID, "pal_gcc_phat_tdoa: need B >= 0, M >= 2, P >= 1");
  if (B == 0) return PAL_OK;
  if (!sig_dev || !pairs_dev || !k_idx_dev || !peak_dev || !gmax_dev || !flags_dev || !ws_dev)
    return fail(PAL_ERR_INVALID, "pal_gcc_phat_tdoa: NULL device pointer");
  if (prm->peak_dist < 1) return fail(PAL_ERR_INVALID, "pal_gcc_phat_tdoa: peak_dist must be >= 1 (scipy: `distance` must be greater or equal to 1)");
  if (prm->num_peaks < 1 || prm->num_peaks > 16) return fail(PAL_ERR_INVALID, "pal_gcc_phat_tdoa: num_peaks must be in 1..16");
  const int n1 = prm->len_first > 0 ? prm->len_first : n_samples;
  const int n2 = prm->len_second > 0 ? prm->len_second : n_samples;
  if (n1 > n_samples || n2 > n_samples) return fail(PAL_ERR_INVALID, "pal_gcc_phat_tdoa: len_first/len_second exceed n_samples");
  if (n1 != n2 && M != 2) return fail(PAL_ERR_INVALID, "pal_gcc_phat_tdoa: unequal lengths need M == 2");
  if (n_samples != kFrame2048 || n1 != kFrame2048 || n2 != kFrame2048)
    return generic_tdoa(sig_dev, B, M, n_samples, n1, n2, pairs_dev, P, prm, k_idx_dev, k_count_dev, peak_dev, gmax_dev,
                        flags_dev, corr_opt_dev, ws_dev, ws_bytes, static_cast<cudaStream_t>(stream_));
  if ((reinterpret_cast<uintptr_t>(sig_dev) & 15u) || (reinterpret_cast<uintptr_t>(ws_dev) & 255u))
    return fail(PAL_ERR_INVALID, "pal_gcc_phat_tdoa: sig_dev must be 16-byte and ws_dev 256-byte aligned");
  if (B * (int64_t)P > 0x7fffffffLL) return fail(PAL_ERR_INVALID, "pal_gcc_phat_tdoa: B*P exceeds 2^31-1; split the batch");

  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  DevInfo di;
  if (int rc = device_info(di)) return rc;

  const size_t per_frame = align_up(size_t(M) * kSpecSlots * sizeof(cpxf), 256);
  const size_t list_bytes = align_up(size_t(B) * P * sizeof(int), 256) + kListTail;
  const size_t hq_frame = size_t(M) * 2 * sizeof(float);
  if (ws_bytes < per_frame + hq_frame + list_bytes + 256) return fail(PAL_ERR_WORKSPACE, "pal_gcc_phat_tdoa: workspace too small");
  char* ws = static_cast<char*>(ws_dev);
  int* list = reinterpret_cast<int*>(ws);
  int* count = reinterpret_cast<int*>(ws + list_bytes - kListTail);
  cpxf* spec = reinterpret_cast<cpxf*>(ws + list_bytes);
  const int64_t chunk = std::min<int64_t>(B, int64_t((ws_bytes - list_bytes - 256) / (per_frame + hq_frame)));
  float* hq = reinterpret_cast<float*>(ws + list_bytes + align_up(size_t(chunk) * per_frame, 256));   // [chunk][M][2]

  const PickParams pp{prm->win_half, prm->peak_dist, prm->thr_method, prm->thr_mult, prm->num_peaks};
  const size_t fwd_smem = sizeof(FwdSmem);
#ifndef PAL_FAST_SMEM_PAD   // tuning experiment: shrink the L1 carve-out
#define PAL_FAST_SMEM_PAD 0
#endif
  const size_t fast_smem = kFastWarps * sizeof(FastWarpSmem) + PAL_FAST_SMEM_PAD;
  const size_t exd_smem = sizeof(ExactSmem<double>);
  PAL_CUDA(cudaFuncSetAttribute(k_pair4095_fast<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fast_smem));
  PAL_CUDA(cudaFuncSetAttribute(k_pair4095_fast<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fast_smem));
  // warp tiles + the tensor-memory slot + (when it fits the opt-in limit) a copy of the pair table
  size_t tmem_smem = kTmemWarps * sizeof(FastWarpSmem) + 16;
#ifndef PAL_PAIRS_IN_SMEM   // A/B measurements: 0 = the kernel always reads the pair table from global memory
#define PAL_PAIRS_IN_SMEM 1
#endif
  const int pairs_in_smem = (PAL_PAIRS_IN_SMEM && tmem_smem + sizeof(int) * 2 * size_t(P) <= di.smem_optin) ? 1 : 0;
  if (pairs_in_smem) tmem_smem += sizeof(int) * 2 * size_t(P);
  // (the cap is the device's opt-in limit whatever P is: callers with different pair counts never shrink it under each other)
  const int tmem_cap = (int)std::max(tmem_smem, di.smem_optin);
  PAL_CUDA(cudaFuncSetAttribute(k_pair4095_tmem<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, tmem_cap));
  PAL_CUDA(cudaFuncSetAttribute(k_pair4095_tmem<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, tmem_cap));
  PAL_CUDA(cudaFuncSetAttribute(k_pair4095_exact<double, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)exd_smem));
  PAL_CUDA(cudaFuncSetAttribute(k_fwd4095, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fwd_smem));

  int fwd_resident = 4, exact_resident = 1;
  PAL_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&fwd_resident, k_fwd4095, kFwdThreads, fwd_smem));
  PAL_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&exact_resident, k_pair4095_exact<double, false>, kExactThreads, exd_smem));
  if (fwd_resident < 1) fwd_resident = 1;
  if (exact_resident < 1) exact_resident = 1;

  if (prm->num_peaks != 1) {
    // num_peaks > 1: every row goes through the exact float64 kernel, straight from the raw
    // frames (not the hot path: main.py:204 always asks for one peak)
    const long long n_items = (long long)B * P;
    const int ge = (int)std::min<long long>(n_items, (long long)di.sms * exact_resident);
    k_pair4095_exact<double, false><<<ge, kExactThreads, exd_smem, stream>>>(
        sig_dev, nullptr, pairs_dev, M, P, n_items, nullptr, nullptr, pp, k_idx_dev, k_count_dev, peak_dev,
        gmax_dev, flags_dev, PAL_FLAG_REFINED, 0u, corr_opt_dev);
    ++g_launches;
    PAL_CUDA(cudaGetLastError());
    return PAL_OK;
  }
  PAL_CUDA(cudaMemsetAsync(count, 0, sizeof(int), stream));

  for (int64_t f0 = 0; f0 < B; f0 += chunk) {
    const int64_t nb = std::min<int64_t>(chunk, B - f0);
    const long long n_items = (long long)nb * P;
    const long long item0 = (long long)f0 * P;
    const float* sig = sig_dev + f0 * M * kFrame2048;
    {
      const long long units = (long long)nb * ((M + 1) / 2);
      const int gf = (int)std::min<long long>(units, (long long)di.sms * fwd_resident);   // persistent blocks: as many as fit
      {
        ProfScope ps(1, stream);
        k_fwd4095<<<gf, kFwdThreads, fwd_smem, stream>>>(sig, M, units, spec, hq);
      }
      ++g_launches;
      const int gp = (int)std::min<long long>((n_items + kFastWarps - 1) / kFastWarps, (long long)di.sms);
      float* corr = corr_opt_dev ? corr_opt_dev + item0 * kN4095 : nullptr;
      {
        ProfScope ps(2, stream);
        if (use_tmem_kernel()) {
          const int gt = (int)std::min<long long>((n_items + kTmemWarps - 1) / kTmemWarps, (long long)di.sms);
          auto kern = corr ? k_pair4095_tmem<true> : k_pair4095_tmem<false>;
          kern<<<gt, kTmemWarps * 32, tmem_smem, stream>>>(spec, hq, pairs_dev, M, P, n_items, pp.win_half, pp.dist,
                                                           prm->tie_eps, k_idx_dev + item0, peak_dev + item0,
                                                           gmax_dev + item0, flags_dev + item0, corr, pairs_in_smem);
        } else {
          auto kern = corr ? k_pair4095_fast<true> : k_pair4095_fast<false>;
          kern<<<gp, kFastWarps * 32, fast_smem, stream>>>(spec, hq, pairs_dev, M, P, n_items, pp.win_half, pp.dist,
                                                           prm->tie_eps, k_idx_dev + item0, peak_dev + item0,
                                                           gmax_dev + item0, flags_dev + item0, corr);
        }
      }
      ++g_launches;
      if (prm->refine) {
        k_compact_flagged<<<(unsigned)((n_items + 255) / 256), 256, 0, stream>>>(flags_dev + item0, n_items, item0,
                                                                                 kRefineMask, list, count);
        ++g_launches;
      }
    }
    PAL_CUDA(cudaGetLastError());
  }
  {
    if (k_count_dev) {
      k_fill_count<<<(unsigned)((B * P + 255) / 256), 256, 0, stream>>>(k_count_dev, B * P);
      ++g_launches;
    }
    if (prm->refine) {
      // float64 re-evaluation of the flagged rows, straight from the raw frames (global item ids)
      ProfScope ps(3, stream);
      k_pair4095_exact<double, false><<<di.sms * exact_resident, kExactThreads, exd_smem, stream>>>(
          sig_dev, nullptr, pairs_dev, M, P, (long long)B * P, list, count, pp, k_idx_dev, nullptr, peak_dev,
          gmax_dev, flags_dev, PAL_FLAG_REFINED, kRefineMask, nullptr);
      ++g_launches;
    }
    PAL_CUDA(cudaGetLastError());
  }
  return PAL_OK;
}

int pal_gcc_phat_tdoa_f64(const double* sig_dev, int64_t B, int32_t M, int32_t n_samples, const int32_t* pairs_dev, int32_t P,
                          const pal_tdoa_params* prm, int32_t* k_idx_dev, int32_t* k_count_dev, float* peak_dev, float* gmax_dev,
                          uint32_t* flags_dev, float* corr_opt_dev, void* ws_dev, size_t ws_bytes, void* stream_) {
  if (!prm) return fail(PAL_ERR_INVALID, "pal_gcc_phat_tdoa_f64: prm is NULL");
  if (B < 0 || M < 2 || P < 1) return fail(PAL_ERR_INVALID, "pal_gcc_phat_tdoa_f64: need B >= 0, M >= 2, P >= 1");
  if (B == 0) return PAL_OK;
  if (!sig_dev || !pairs_dev || !k_idx_dev || !peak_dev || !gmax_dev || !flags_dev || !ws_dev)
    return fail(PAL_ERR_INVALID, "pal_gcc_phat_tdoa_f64: NULL device pointer");
  if (prm->peak_dist < 1) return fail(PAL_ERR_INVALID, "pal_gcc_phat_tdoa_f64: peak_dist must be >= 1 (scipy: `distance` must be greater or equal to 1)");
  if (prm->num_peaks < 1 || prm->num_peaks > 16) return fail(PAL_ERR_INVALID, "pal_gcc_phat_tdoa_f64: num_peaks must be in 1..16");
  const int n1 = prm->len_first > 0 ? prm->len_first : n_samples;
  const int n2 = prm->len_second > 0 ? prm->len_second : n_samples;
  if (n1 > n_samples || n2 > n_samples) return fail(PAL_ERR_INVALID, "pal_gcc_phat_tdoa_f64: len_first/len_second exceed n_samples");
  if (n1 != n2 && M != 2) return fail(PAL_ERR_INVALID, "pal_gcc_phat_tdoa_f64: unequal lengths need M == 2");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (n_samples != kFrame2048 || n1 != kFrame2048 || n2 != kFrame2048)
    return generic_tdoa(nullptr, B, M, n_samples, n1, n2, pairs_dev, P, prm, k_idx_dev, k_count_dev, peak_dev, gmax_dev, flags_dev,
                        corr_opt_dev, ws_dev, ws_bytes, stream, sig_dev);
  if (B * (int64_t)P > 0x7fffffffLL) return fail(PAL_ERR_INVALID, "pal_gcc_phat_tdoa_f64: B*P exceeds 2^31-1; split the batch");
  DevInfo di;
  if (int rc = device_info(di)) return rc;
  const PickParams pp{prm->win_half, prm->peak_dist, prm->thr_method, prm->thr_mult, prm->num_peaks};
  const size_t exd_smem = sizeof(ExactSmem<double>);
  PAL_CUDA(cudaFuncSetAttribute(k_pair4095_exact<double, false, double>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)exd_smem));
  int resident = 1;
  PAL_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&resident, k_pair4095_exact<double, false, double>, kExactThreads, exd_smem));
  const long long n_items = (long long)B * P;
  const int ge = (int)std::min<long long>(n_items, (long long)di.sms * std::max(resident, 1));
  k_pair4095_exact<double, false, double><<<ge, kExactThreads, exd_smem, stream>>>(
      sig_dev, nullptr, pairs_dev, M, P, n_items, nullptr, nullptr, pp, k_idx_dev, k_count_dev, peak_dev, gmax_dev, flags_dev,
      PAL_FLAG_REFINED, 0u, corr_opt_dev);
  ++g_launches;
  PAL_CUDA(cudaGetLastError());
  return PAL_OK;
}

int pal_pcm16_to_f32(const int16_t* in_dev, int64_t count, float scale, float* out_dev, void* stream_) {
  if (count < 0 || (count > 0 && (!in_dev || !out_dev))) return fail(PAL_ERR_INVALID, "pal_pcm16_to_f32: bad argument");
  if ((reinterpret_cast<uintptr_t>(in_dev) & 1u) || (reinterpret_cast<uintptr_t>(out_dev) & 3u))
    return fail(PAL_ERR_INVALID, "pal_pcm16_to_f32: misaligned pointer");
  if (count == 0) return PAL_OK;
  DevInfo di;
  if (int rc = device_info(di)) return rc;
  const int head = int(((16u - (reinterpret_cast<uintptr_t>(in_dev) & 15u)) & 15u) / 2u);
  const unsigned grid = (unsigned)std::min<long long>((count / 8 + 255) / 256 + 1, 16LL * di.sms);
  k_pcm16_to_f32<<<grid, 256, 0, static_cast<cudaStream_t>(stream_)>>>(reinterpret_cast<const short*>(in_dev), count, scale,
                                                                       out_dev, head);
  ++g_launches;
  PAL_CUDA(cudaGetLastError());
  return PAL_OK;
}

int pal_tdoa_seconds(const int32_t* k_idx_dev, int64_t count, int32_t n_second, double fs, double* out_dev, void* stream_) {
  if (count < 0 || n_second < 1 || !(fs > 0.0) || (count > 0 && (!k_idx_dev || !out_dev)))
    return fail(PAL_ERR_INVALID, "pal_tdoa_seconds: bad argument");
  if (count == 0) return PAL_OK;
  k_tdoa_seconds<<<(unsigned)((count + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream_)>>>(k_idx_dev, count, n_second - 1, fs,
                                                                                                   out_dev);
  ++g_launches;
  PAL_CUDA(cudaGetLastError());
  return PAL_OK;
}

int pal_image_sources_workspace(int32_t n_planes, int32_t k_max, int64_t n_scenes, size_t* bytes) {
  if (n_planes < 0 || k_max < 1 || n_scenes < 0 || !bytes) return fail(PAL_ERR_INVALID, "pal_image_sources_workspace: bad argument");
  int sms = 148;
  DevInfo di;
  if (device_info(di) == PAL_OK && di.sms > 0) sms = di.sms;
  *bytes = palhost::image_scratch_per_block(std::max(n_planes, 1), k_max) * size_t(palhost::image_grid(std::max<int64_t>(n_scenes, 1), sms));
  return PAL_OK;
}

int pal_image_sources(const double* sources_dev, int64_t n_scenes, const double* planes_dev, int64_t plane_stride,
                      const int32_t* plane_mat_dev, int32_t n_planes, const double* mat_abs_dev, const double* mat_freq_dev, const double* mics_dev,
                      int32_t n_mics, int64_t mic_stride, int32_t max_order, double frequency, double threshold,
                      int32_t round_decimals, int32_t k_max, double* out_pos_dev, int32_t* out_mat_dev,
                      int32_t* out_count_dev, void* ws_dev, size_t ws_bytes, void* stream_) {
  if (n_scenes < 0 || n_planes < 0 || n_mics < 1 || k_max < 1 || max_order < 0)
    return fail(PAL_ERR_INVALID, "pal_image_sources: bad size argument");
  if (n_scenes == 0) return PAL_OK;
  if (!sources_dev || !mics_dev || !out_pos_dev || !out_mat_dev || !out_count_dev || !ws_dev ||
      (n_planes > 0 && (!planes_dev || !plane_mat_dev || !mat_abs_dev || !mat_freq_dev)))
    return fail(PAL_ERR_INVALID, "pal_image_sources: NULL device pointer");
  DevInfo di;
  if (int rc = device_info(di)) return rc;
  const size_t per_block = palhost::image_scratch_per_block(std::max(n_planes, 1), k_max);
  const int grid = palhost::image_grid(n_scenes, di.sms);
  if (ws_bytes < per_block * size_t(grid)) return fail(PAL_ERR_WORKSPACE, "pal_image_sources: workspace too small");
  double scale = 1.0;
  for (int i = 0; i < round_decimals; ++i) scale *= 10.0;
  for (int i = 0; i > round_decimals; --i) scale /= 10.0;
  ImgParams ip{n_planes, n_mics, max_order, k_max, frequency, threshold, scale};
  palhost::k_image_sources<<<grid, palhost::kImgThreads, 64 * sizeof(int), static_cast<cudaStream_t>(stream_)>>>(
      ip, sources_dev, n_scenes, planes_dev, plane_stride, plane_mat_dev, mat_abs_dev, mat_freq_dev, mics_dev, mic_stride, out_pos_dev,
      out_mat_dev, out_count_dev, static_cast<char*>(ws_dev), per_block);
  ++g_launches;
  PAL_CUDA(cudaGetLastError());
  return PAL_OK;
}

int pal_path_table(const double* source_dev, const double* img_pos_dev, const int32_t* img_mat_dev, int32_t n_img,
                   const double* mics_dev, int32_t n_mics, const double* mat_abs_dev, const double* mat_freq_dev,
                   int32_t air_mat, double frequency, double c_sound, double* tau_dev, double* gain_dev, void* stream_) {
  if (n_img < 0 || n_mics < 1 || !source_dev || !mics_dev || !mat_abs_dev || !mat_freq_dev || !tau_dev || !gain_dev ||
      (n_img > 0 && (!img_pos_dev || !img_mat_dev)))
    return fail(PAL_ERR_INVALID, "pal_path_table: bad argument");
  const long long total = (long long)n_mics * (n_img + 1);
  palhost::k_path_table<<<(unsigned)((total + 127) / 128), 128, 0, static_cast<cudaStream_t>(stream_)>>>(
      source_dev, img_pos_dev, img_mat_dev, n_img, mics_dev, n_mics, mat_abs_dev, mat_freq_dev, air_mat, frequency, c_sound,
      tau_dev, gain_dev);
  ++g_launches;
  PAL_CUDA(cudaGetLastError());
  return PAL_OK;
}

int pal_render_workspace(int32_t N, int32_t n_mics, size_t* bytes, size_t* min_bytes) {
  if (N < 2 || n_mics < 1 || !bytes) return fail(PAL_ERR_INVALID, "pal_render_workspace: bad argument");
  *bytes = palhost::render_full_bytes(N, n_mics);
  if (min_bytes) *min_bytes = palhost::render_min_bytes(N, n_mics);
  return PAL_OK;
}

int pal_render_scene(const float* base_dev, int32_t n_base, int32_t N, const double* tau_dev, const double* gain_dev,
                     int32_t n_mics, int32_t n_paths, double fs, int32_t n_keep, int32_t flags, float* out_dev,
                     void* ws_dev, size_t ws_bytes, void* stream_) {
  if (!base_dev || !tau_dev || !gain_dev || !out_dev || !ws_dev) return fail(PAL_ERR_INVALID, "pal_render_scene: NULL device pointer");
  if (n_base < 1 || N < n_base || n_mics < 1 || n_paths < 1 || n_keep < 1 || n_keep > N)
    return fail(PAL_ERR_INVALID, "pal_render_scene: need 1 <= n_base <= N, 1 <= n_keep <= N, n_mics, n_paths >= 1");
  if (int(0.01 * N) < 1)
    return fail(PAL_ERR_INVALID, "pal_render_scene: N < 100 (the reference's fade window of int(0.01*N) samples is empty and numpy cannot broadcast it)");
  if (reinterpret_cast<uintptr_t>(ws_dev) & 255u) return fail(PAL_ERR_INVALID, "pal_render_scene: ws_dev must be 256-byte aligned");
  DevInfo di;
  if (int rc = device_info(di)) return rc;
  if (ws_bytes < palhost::render_min_bytes(N, n_mics)) return fail(PAL_ERR_WORKSPACE, "pal_render_scene: workspace too small");
  cudaStream_t st = static_cast<cudaStream_t>(stream_);
  const RenderRows rr{tau_dev, gain_dev, nullptr, nullptr, n_paths, n_mics};
  cudaError_t e = palhost::render_rows(base_dev, n_base, N, rr, n_mics, fs, n_keep, out_dev, static_cast<char*>(ws_dev), ws_bytes,
                                       st, di.sms);
  if (e == cudaSuccess && (flags & PAL_RENDER_NORMALISE_COMPRESS)) e = palhost::normalise_rows(out_dev, n_mics, n_keep, st, di.sms);
  if (e != cudaSuccess) return cuda_fail(e, "pal_render_scene");
  return PAL_OK;
}

int pal_path_table_batched(const double* sources_dev, const double* img_pos_dev, const int32_t* img_mat_dev,
                           const int32_t* img_count_dev, int64_t n_scenes, int32_t k_max, const double* mics_dev,
                           int32_t n_mics, int64_t mic_stride, const double* mat_abs_dev, const double* mat_freq_dev,
                           int32_t air_mat, double frequency, double c_sound, int32_t k_stride, double* tau_dev,
                           double* gain_dev, int32_t* path_count_dev, double* max_tau_dev, void* stream_) {
  if (n_scenes < 0 || k_max < 1 || n_mics < 1 || k_stride < k_max + 1)
    return fail(PAL_ERR_INVALID, "pal_path_table_batched: need n_scenes >= 0, k_max >= 1, n_mics >= 1, k_stride >= k_max + 1");
  if (n_scenes == 0) return PAL_OK;
  if (!sources_dev || !img_pos_dev || !img_mat_dev || !img_count_dev || !mics_dev || !mat_abs_dev || !mat_freq_dev || !tau_dev ||
      !gain_dev || !path_count_dev || !max_tau_dev)
    return fail(PAL_ERR_INVALID, "pal_path_table_batched: NULL device pointer");
  DevInfo di;
  if (int rc = device_info(di)) return rc;
  palhost::k_path_table_batched<<<(unsigned)std::min<long long>(n_scenes, 16LL * di.sms), 128, 64, static_cast<cudaStream_t>(stream_)>>>(
      sources_dev, img_pos_dev, img_mat_dev, img_count_dev, n_scenes, k_max, mics_dev, n_mics, mic_stride, mat_abs_dev,
      mat_freq_dev, air_mat, frequency, c_sound, k_stride, tau_dev, gain_dev, path_count_dev, max_tau_dev);
  ++g_launches;
  PAL_CUDA(cudaGetLastError());
  return PAL_OK;
}

int pal_render_scenes_workspace(int32_t N, int64_t n_rows, size_t* bytes, size_t* min_bytes) {
  if (N < 2 || n_rows < 1 || !bytes) return fail(PAL_ERR_INVALID, "pal_render_scenes_workspace: bad argument");
  *bytes = palhost::render_full_bytes(N, n_rows);
  if (min_bytes) *min_bytes = palhost::render_min_bytes(N, 1);
  return PAL_OK;
}

int pal_render_scenes(const float* base_dev, int32_t n_base, int32_t N, const double* tau_dev, const double* gain_dev,
                      const int32_t* path_count_dev, int32_t k_stride, const int64_t* scene_index_dev, int64_t n_bucket_scenes,
                      int32_t n_mics, double fs, int32_t n_keep, float* out_dev, void* ws_dev, size_t ws_bytes, void* stream_) {
  if (!base_dev || !tau_dev || !gain_dev || !path_count_dev || !out_dev || !ws_dev)
    return fail(PAL_ERR_INVALID, "pal_render_scenes: NULL device pointer");
  if (n_base < 1 || N < n_base || n_mics < 1 || k_stride < 1 || n_keep < 1 || n_keep > N || n_bucket_scenes < 0)
    return fail(PAL_ERR_INVALID, "pal_render_scenes: need 1 <= n_base <= N, 1 <= n_keep <= N, n_mics, k_stride >= 1");
  if (int(0.01 * N) < 1) return fail(PAL_ERR_INVALID, "pal_render_scenes: N < 100 (empty fade window, signal_processing.py:74-79)");
  if (reinterpret_cast<uintptr_t>(ws_dev) & 255u) return fail(PAL_ERR_INVALID, "pal_render_scenes: ws_dev must be 256-byte aligned");
  if (n_bucket_scenes == 0) return PAL_OK;
  DevInfo di;
  if (int rc = device_info(di)) return rc;
  if (ws_bytes < palhost::render_min_bytes(N, 1)) return fail(PAL_ERR_WORKSPACE, "pal_render_scenes: workspace too small");
  const RenderRows rr{tau_dev, gain_dev, path_count_dev, reinterpret_cast<const long long*>(scene_index_dev), k_stride, n_mics};
  cudaError_t e = palhost::render_rows(base_dev, n_base, N, rr, n_bucket_scenes * n_mics, fs, n_keep, out_dev,
                                       static_cast<char*>(ws_dev), ws_bytes, static_cast<cudaStream_t>(stream_), di.sms);
  if (e != cudaSuccess) return cuda_fail(e, "pal_render_scenes");
  return PAL_OK;
}

int pal_render_plan_bytes(int32_t N, size_t* plan_bytes, size_t* scratch_bytes) {
  if (N < 100 || N > (1 << 27) || !plan_bytes) return fail(PAL_ERR_INVALID, "pal_render_plan_bytes: bad argument");
  *plan_bytes = palhost::render_plan_bytes(N);
  if (scratch_bytes) *scratch_bytes = palhost::render_plan_scratch_bytes(N);
  return PAL_OK;
}

int pal_render_plan(const float* base_dev, int32_t n_base, int32_t N, void* plan_dev, size_t plan_bytes, void* scratch_dev,
                    size_t scratch_bytes, void* stream_) {
  if (!base_dev || !plan_dev || !scratch_dev) return fail(PAL_ERR_INVALID, "pal_render_plan: NULL device pointer");
  if (n_base < 1 || N < n_base || int(0.01 * N) < 1) return fail(PAL_ERR_INVALID, "pal_render_plan: need 1 <= n_base <= N, N >= 100");
  if ((reinterpret_cast<uintptr_t>(plan_dev) | reinterpret_cast<uintptr_t>(scratch_dev)) & 255u)
    return fail(PAL_ERR_INVALID, "pal_render_plan: plan_dev and scratch_dev must be 256-byte aligned");
  if (plan_bytes < palhost::render_plan_bytes(N) || scratch_bytes < palhost::render_plan_scratch_bytes(N))
    return fail(PAL_ERR_WORKSPACE, "pal_render_plan: plan or scratch memory too small");
  DevInfo di;
  if (int rc = device_info(di)) return rc;
  cudaError_t e = palhost::build_render_plan(base_dev, n_base, N, static_cast<char*>(plan_dev), static_cast<cpxf*>(scratch_dev),
                                             static_cast<cudaStream_t>(stream_), di.sms);
  if (e != cudaSuccess) return cuda_fail(e, "pal_render_plan");
  return PAL_OK;
}

int pal_render_rows_workspace(int32_t N, int64_t n_rows, size_t* bytes, size_t* min_bytes) {
  if (N < 1 || n_rows < 1 || !bytes) return fail(PAL_ERR_INVALID, "pal_render_rows_workspace: bad argument");
  *bytes = size_t(std::min<long long>(n_rows, 4096)) * palhost::render_row_bytes(N) + 1024;
  if (min_bytes) *min_bytes = palhost::render_rows_min_bytes(N);
  return PAL_OK;
}

int pal_render_scenes_planned(const void* plan_dev, size_t plan_bytes, int32_t n_base, int32_t N, const double* tau_dev,
                              const double* gain_dev, const int32_t* path_count_dev, int32_t k_stride,
                              const int64_t* scene_index_dev, int64_t n_bucket_scenes, int32_t n_mics, double fs,
                              int32_t n_keep, float* out_dev, void* ws_dev, size_t ws_bytes, void* stream_) {
  if (!plan_dev || !tau_dev || !gain_dev || !path_count_dev || !out_dev || !ws_dev)
    return fail(PAL_ERR_INVALID, "pal_render_scenes_planned: NULL device pointer");
  if (n_base < 1 || N < n_base || n_mics < 1 || k_stride < 1 || n_keep < 1 || n_keep > N || n_bucket_scenes < 0)
    return fail(PAL_ERR_INVALID, "pal_render_scenes_planned: need 1 <= n_base <= N, 1 <= n_keep <= N, n_mics, k_stride >= 1");
  if (int(0.01 * N) < 1) return fail(PAL_ERR_INVALID, "pal_render_scenes_planned: N < 100 (empty fade window, signal_processing.py:74-79)");
  if ((reinterpret_cast<uintptr_t>(ws_dev) | reinterpret_cast<uintptr_t>(plan_dev)) & 255u)
    return fail(PAL_ERR_INVALID, "pal_render_scenes_planned: plan_dev and ws_dev must be 256-byte aligned");
  if (plan_bytes < palhost::render_plan_bytes(N)) return fail(PAL_ERR_WORKSPACE, "pal_render_scenes_planned: plan memory too small for N");
  if (n_bucket_scenes == 0) return PAL_OK;
  DevInfo di;
  if (int rc = device_info(di)) return rc;
  if (ws_bytes < palhost::render_rows_min_bytes(N)) return fail(PAL_ERR_WORKSPACE, "pal_render_scenes_planned: workspace too small");
  const RenderRows rr{tau_dev, gain_dev, path_count_dev, reinterpret_cast<const long long*>(scene_index_dev), k_stride, n_mics};
  const palhost::RenderPlan rp = palhost::carve_render_plan(N, static_cast<char*>(const_cast<void*>(plan_dev)));
  cudaError_t e = palhost::render_rows_planned(rp, N, rr, n_bucket_scenes * n_mics, fs, n_keep, out_dev, static_cast<char*>(ws_dev),
                                               ws_bytes, static_cast<cudaStream_t>(stream_), di.sms);
  if (e != cudaSuccess) return cuda_fail(e, "pal_render_scenes_planned");
  return PAL_OK;
}

int pal_render_scenes_grouped(int32_t n_buckets, const void* const* plan_dev_of_bucket, const int32_t* N_of_bucket,
                              const int64_t* first_scene_of_bucket, const double* tau_dev, const double* gain_dev,
                              const int32_t* path_count_dev, int32_t k_stride, const int64_t* scene_index_dev, int32_t n_mics,
                              double fs, int32_t n_keep, float* out_dev, void* ws_dev, size_t ws_bytes, void* stream_) {
  if (n_buckets < 0 || n_mics < 1 || k_stride < 1 || n_keep < 1) return fail(PAL_ERR_INVALID, "pal_render_scenes_grouped: bad size argument");
  if (n_buckets == 0) return PAL_OK;
  if (!plan_dev_of_bucket || !N_of_bucket || !first_scene_of_bucket || !tau_dev || !gain_dev || !path_count_dev || !scene_index_dev ||
      !out_dev || !ws_dev)
    return fail(PAL_ERR_INVALID, "pal_render_scenes_grouped: NULL pointer");
  if (reinterpret_cast<uintptr_t>(ws_dev) & 255u) return fail(PAL_ERR_INVALID, "pal_render_scenes_grouped: ws_dev must be 256-byte aligned");
  DevInfo di;
  if (int rc = device_info(di)) return rc;
  std::vector<palhost::GroupedBucketIn> in(n_buckets);
  for (int i = 0; i < n_buckets; ++i) {
    const int N = N_of_bucket[i];
    const long long cnt = first_scene_of_bucket[i + 1] - first_scene_of_bucket[i];
    if (!plan_dev_of_bucket[i] || (reinterpret_cast<uintptr_t>(plan_dev_of_bucket[i]) & 255u) || N < 100 || n_keep > N || cnt < 1)
      return fail(PAL_ERR_INVALID, "pal_render_scenes_grouped: bad bucket (NULL / misaligned plan, N < 100, n_keep > N or no scene)");
    in[i] = palhost::GroupedBucketIn{static_cast<const char*>(plan_dev_of_bucket[i]), N, first_scene_of_bucket[i], cnt};
  }
  const RenderRows rr{tau_dev, gain_dev, path_count_dev, nullptr, k_stride, n_mics};
  cudaError_t e = palhost::render_grouped(in.data(), n_buckets, rr, reinterpret_cast<const long long*>(scene_index_dev), fs, n_keep,
                                          out_dev, static_cast<char*>(ws_dev), ws_bytes, static_cast<cudaStream_t>(stream_), di.sms);
  if (e == cudaErrorNotSupported) return fail(PAL_ERR_UNSUPPORTED, "pal_render_scenes_grouped: a bucket has no compile-time plan or exceeds the workspace; use pal_render_scenes_planned per bucket");
  if (e != cudaSuccess) return cuda_fail(e, "pal_render_scenes_grouped");
  return PAL_OK;
}

int pal_normalise_compress(float* rows_dev, int64_t n_rows, int32_t n, float threshold, float epsilon, int32_t mode,
                           void* stream_) {
  if (n_rows < 0 || n < 1 || (n_rows > 0 && !rows_dev) || (mode != 0 && mode != 1))
    return fail(PAL_ERR_INVALID, "pal_normalise_compress: bad argument");
  if (n_rows == 0) return PAL_OK;
  DevInfo di;
  if (int rc = device_info(di)) return rc;
  palhost::k_normalise_compress<<<(unsigned)std::min<long long>(n_rows, 8LL * di.sms), palhost::kGT, 64,
                                  static_cast<cudaStream_t>(stream_)>>>(rows_dev, n_rows, n, threshold, epsilon, mode == 1);
  ++g_launches;
  PAL_CUDA(cudaGetLastError());
  return PAL_OK;
}

int pal_filtfilt_workspace(int64_t n_rows, int32_t n, int32_t padlen, size_t* bytes) {
  if (n_rows < 0 || n < 1 || padlen < 0 || !bytes) return fail(PAL_ERR_INVALID, "pal_filtfilt_workspace: bad argument");
  *bytes = size_t((n_rows + 31) / 32) * size_t(n + 2 * padlen) * 32 * sizeof(double) + 256;
  return PAL_OK;
}

int pal_filtfilt(const void* x_dev, int64_t n_rows, int32_t n, int32_t io_f32, const double* b, const double* a,
                 const double* zi, int32_t ntaps, int32_t padlen, void* y_dev, void* ws_dev, size_t ws_bytes, void* stream_) {
  if (n_rows < 0 || n < 1 || !b || !a || !zi) return fail(PAL_ERR_INVALID, "pal_filtfilt: bad argument");
  if (ntaps < 2 || ntaps > kFiltMaxTaps) return fail(PAL_ERR_UNSUPPORTED, "pal_filtfilt: need 2 <= ntaps <= 16");
  if (padlen < 0 || n <= padlen)
    return fail(PAL_ERR_INVALID, "pal_filtfilt: The length of the input vector x must be greater than padlen");
  if (a[0] != 1.0) return fail(PAL_ERR_INVALID, "pal_filtfilt: a[0] must be 1 (normalise b and a by a[0] first)");
  if (n_rows == 0) return PAL_OK;
  if (!x_dev || !y_dev || !ws_dev) return fail(PAL_ERR_INVALID, "pal_filtfilt: NULL device pointer");
  size_t need = 0;
  pal_filtfilt_workspace(n_rows, n, padlen, &need);
  if (ws_bytes < need - 256) return fail(PAL_ERR_WORKSPACE, "pal_filtfilt: workspace too small");
  DevInfo di;
  if (int rc = device_info(di)) return rc;
  FiltParams fp{};
  fp.ntaps = ntaps;
  fp.padlen = padlen;
  for (int i = 0; i < kFiltMaxTaps; ++i) {
    fp.b[i] = i < ntaps ? b[i] : 0.0;
    fp.a[i] = i < ntaps ? a[i] : 0.0;
    fp.zi[i] = i < ntaps - 1 ? zi[i] : 0.0;
  }
  const long long groups = (n_rows + 31) / 32;
  const int wpb = kFiltThreads / 32;
  const unsigned grid = (unsigned)std::min<long long>((groups + wpb - 1) / wpb, 16LL * di.sms);
  const size_t smem = size_t(wpb) * 32 * 33 * sizeof(double);
  cudaStream_t st = static_cast<cudaStream_t>(stream_);
  if (io_f32)
    k_filtfilt<float><<<grid, kFiltThreads, smem, st>>>(static_cast<const float*>(x_dev), n_rows, n, fp, static_cast<double*>(ws_dev),
                                                        static_cast<float*>(y_dev));
  else
    k_filtfilt<double><<<grid, kFiltThreads, smem, st>>>(static_cast<const double*>(x_dev), n_rows, n, fp,
                                                         static_cast<double*>(ws_dev), static_cast<double*>(y_dev));
  ++g_launches;
  PAL_CUDA(cudaGetLastError());
  return PAL_OK;
}

int pal_sync_align_workspace(int64_t n_scenes, int32_t n_ch, int32_t ld, size_t* bytes, size_t* min_bytes) {
  if (n_scenes < 0 || n_ch < 1 || ld < 2 || ld > (1 << 28) || !bytes) return fail(PAL_ERR_INVALID, "pal_sync_align_workspace: bad argument");
  *bytes = palhost::sync_full_bytes<double>(ld, n_scenes, n_ch);
  if (min_bytes) *min_bytes = palhost::sync_min_bytes<double>(ld, n_ch);
  return PAL_OK;
}

int pal_sync_align(const double* sig_dev, int64_t n_scenes, int32_t n_ch, int32_t ld, const int32_t* lens_dev,
                   int32_t* ref_idx_dev, int32_t* peak_index_dev, double* absmax_dev, double* win_dev,
                   double* energy_dev, void* ws_dev, size_t ws_bytes, void* stream_) {
  if (n_scenes < 0 || n_ch < 1 || ld < 2 || ld > (1 << 28)) return fail(PAL_ERR_INVALID, "pal_sync_align: bad size argument");
  if (n_scenes == 0) return PAL_OK;
  if (n_scenes * (int64_t)n_ch > 0x3fffffffLL) return fail(PAL_ERR_INVALID, "pal_sync_align: too many rows; split the batch");
  if (!sig_dev || !ref_idx_dev || !peak_index_dev || !absmax_dev || !win_dev || !ws_dev)
    return fail(PAL_ERR_INVALID, "pal_sync_align: NULL device pointer");
  if (reinterpret_cast<uintptr_t>(ws_dev) & 255u) return fail(PAL_ERR_INVALID, "pal_sync_align: ws_dev must be 256-byte aligned");
  if (ws_bytes < palhost::sync_min_bytes<double>(ld, n_ch)) return fail(PAL_ERR_WORKSPACE, "pal_sync_align: workspace too small");
  DevInfo di;
  if (int rc = device_info(di)) return rc;
  palhost::SyncCall c{sig_dev, (long long)n_scenes, n_ch, ld, lens_dev, ref_idx_dev, peak_index_dev, absmax_dev, win_dev,
                      energy_dev, static_cast<cudaStream_t>(stream_), di.sms};
  cudaError_t e = palhost::run_sync_align<double>(c, static_cast<char*>(ws_dev), ws_bytes);
  if (e != cudaSuccess) return cuda_fail(e, "pal_sync_align");
  return PAL_OK;
}

int pal_pad_rows(const void* in_dev, int64_t n_rows, int64_t ld_in, const int32_t* lens_dev, const int32_t* pad_left_dev,
                 void* out_dev, int64_t ld_out, int32_t io_f32, void* stream_) {
  if (n_rows < 0 || ld_in < 1 || ld_out < 1 || ld_in > 0x7fffffffLL) return fail(PAL_ERR_INVALID, "pal_pad_rows: bad size argument");
  if (n_rows == 0) return PAL_OK;
  if (!in_dev || !out_dev || !pad_left_dev) return fail(PAL_ERR_INVALID, "pal_pad_rows: NULL device pointer");
  DevInfo di;
  if (int rc = device_info(di)) return rc;
  const long long total = (long long)n_rows * ld_out;
  const unsigned grid = (unsigned)std::min<long long>((total + 255) / 256, 32LL * di.sms);
  cudaStream_t st = static_cast<cudaStream_t>(stream_);
  if (io_f32)
    palhost::k_pad_rows<float><<<grid, 256, 0, st>>>(static_cast<const float*>(in_dev), n_rows, ld_in, (int)ld_in, lens_dev,
                                                     pad_left_dev, static_cast<float*>(out_dev), ld_out);
  else
    palhost::k_pad_rows<double><<<grid, 256, 0, st>>>(static_cast<const double*>(in_dev), n_rows, ld_in, (int)ld_in, lens_dev,
                                                      pad_left_dev, static_cast<double*>(out_dev), ld_out);
  ++g_launches;
  PAL_CUDA(cudaGetLastError());
  return PAL_OK;
}

int pal_solve_positions_workspace(int32_t n_pairs, size_t* bytes) {
  if (n_pairs < 1 || !bytes) return fail(PAL_ERR_INVALID, "pal_solve_positions_workspace: bad argument");
  int sms = 148;
  DevInfo di;
  if (device_info(di) == PAL_OK && di.sms > 0) sms = di.sms;
  *bytes = size_t(solve_grid(sms)) * (kSolveThreads / 32) * size_t(n_pairs) * sizeof(double) + 256;
  return PAL_OK;
}

int pal_solve_positions(const double* mics_dev, int64_t mic_stride, int32_t n_mics, const int32_t* pairs_dev, int32_t n_pairs,
                        const double* tdoa_dev, const double* weights_dev, const double* x0_dev, const double* lo_dev,
                        const double* hi_dev, int64_t n_scenes, double c_sound, double buffer, int32_t max_iter, double xtol,
                        double ftol, double gtol, double* out_pos_dev, double* out_cost_dev, int32_t* out_iter_dev, void* ws_dev,
                        size_t ws_bytes, void* stream_) {
  if (n_scenes < 0 || n_mics < 2 || n_pairs < 1 || max_iter < 1 || !(c_sound > 0.0) || mic_stride < 0)
    return fail(PAL_ERR_INVALID, "pal_solve_positions: need n_scenes >= 0, n_mics >= 2, n_pairs >= 1, max_iter >= 1, c_sound > 0");
  if ((lo_dev == nullptr) != (hi_dev == nullptr)) return fail(PAL_ERR_INVALID, "pal_solve_positions: lo_dev and hi_dev go together");
  if (n_scenes == 0) return PAL_OK;
  if (!mics_dev || !pairs_dev || !tdoa_dev || !out_pos_dev || !ws_dev) return fail(PAL_ERR_INVALID, "pal_solve_positions: NULL device pointer");
  DevInfo di;
  if (int rc = device_info(di)) return rc;
  size_t need = 0;
  pal_solve_positions_workspace(n_pairs, &need);
  if (ws_bytes < need - 256) return fail(PAL_ERR_WORKSPACE, "pal_solve_positions: workspace too small");
  const SolveParams sp{n_mics, n_pairs, max_iter, c_sound, buffer, xtol, ftol, gtol};
  const int wpb = kSolveThreads / 32;
  const unsigned grid = (unsigned)std::min<long long>((n_scenes + wpb - 1) / wpb, solve_grid(di.sms));
  k_solve_positions<<<grid, kSolveThreads, 0, static_cast<cudaStream_t>(stream_)>>>(
      sp, mics_dev, mic_stride, pairs_dev, tdoa_dev, weights_dev, x0_dev, lo_dev, hi_dev, n_scenes, static_cast<double*>(ws_dev),
      out_pos_dev, out_cost_dev, out_iter_dev);
  ++g_launches;
  PAL_CUDA(cudaGetLastError());
  return PAL_OK;
}

}  // extern "C"
