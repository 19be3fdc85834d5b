"""ctypes bindings of the host-emulated kernels (tests only)."""
import ctypes as C

import numpy as np

from .build_emu import build

_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
    return _lib


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t)) if a is not None else None


def pairs_of(m):
    return np.array([(i, j) for i in range(m) for j in range(i + 1, m)], dtype=np.int32)


class Spectra(np.ndarray):
    """[B, M, 2080, 2] whitened spectra plus .hq [B, M, 2]: per channel the whitening bound h and the rounding-noise term q
    (pal_pfa4095.cuh: whiten_bin)."""
    hq = None


def fwd4095(sig):
    b, m, n = sig.shape
    assert n == 2048 and sig.dtype == np.float32
    spec = np.zeros((b, m, 2080, 2), np.float32).view(Spectra)
    spec.hq = np.zeros((b, m, 2), np.float32)
    lib().emu_fwd4095(_p(sig, C.c_float), m, C.c_longlong(b), _p(spec, C.c_float), _p(spec.hq, C.c_float), 2)
    return spec


def pair_fast(spec, pairs, win_half, dist, eps=2e-6, want_corr=False, phase_sync=True):
    b, m = spec.shape[:2]
    p = len(pairs)
    k = np.zeros((b, p), np.int32)
    pk = np.zeros((b, p), np.float32)
    gm = np.zeros((b, p), np.float32)
    fl = np.zeros((b, p), np.uint32)
    corr = np.zeros((b, p, 4095), np.float32) if want_corr else None
    hq = getattr(spec, "hq", None)
    hq = np.zeros((b, m, 2), np.float32) if hq is None else np.ascontiguousarray(hq, np.float32)
    lib().emu_pair4095_fast(_p(spec, C.c_float), _p(hq, C.c_float), _p(pairs, C.c_int), m, p, C.c_longlong(b), win_half, dist,
                            C.c_float(eps), _p(k, C.c_int), _p(pk, C.c_float), _p(gm, C.c_float),
                            _p(fl, C.c_uint), _p(corr, C.c_float), 3, int(phase_sync))
    return k, pk, gm, fl, corr


def pair_exact(mode, sig, spec, pairs, win_half, dist, method=0, mult=1.0, num_peaks=1, items=None,
               want_corr=False):
    b, m = sig.shape[:2]
    p = len(pairs)
    k = np.full((b, p, num_peaks), -7, np.int32)
    cnt = np.zeros((b, p), np.int32)
    pk = np.zeros((b, p), np.float32)
    gm = np.zeros((b, p), np.float32)
    fl = np.zeros((b, p), np.uint32)
    corr = np.zeros((b, p, 4095), np.float32) if want_corr else None
    il = ic = None
    if items is not None:
        il = np.asarray(items, np.int32)
        ic = np.array([len(il)], np.int32)
    lib().emu_pair4095_exact(mode, _p(sig, C.c_float), _p(spec, C.c_float), _p(pairs, C.c_int), m, p,
                             C.c_longlong(b), _p(il, C.c_int), _p(ic, C.c_int), win_half, dist, method,
                             C.c_float(mult), num_peaks, _p(k, C.c_int), _p(cnt, C.c_int), _p(pk, C.c_float),
                             _p(gm, C.c_float), _p(fl, C.c_uint), _p(corr, C.c_float), 2)
    return k, cnt, pk, gm, fl, corr


def peakpick_f64(c, c0, win_half, dist, method=0, mult=1.0, num_peaks=1):
    c = np.ascontiguousarray(c, np.float64)
    out = np.zeros(16, np.int32)
    cnt = np.zeros(1, np.int32)
    fl = np.zeros(1, np.uint32)
    lib().emu_peakpick_f64(_p(c, C.c_double), len(c), c0, win_half, dist, method, C.c_double(mult), num_peaks,
                           _p(out, C.c_int), _p(cnt, C.c_int), _p(fl, C.c_uint))
    return list(out[: cnt[0]]), int(fl[0])


def generic_gcc_phat(sig, n1, n2, pairs, win_half, dist, method=0, mult=1.0, num_peaks=1, eps=0.0, use_double=False):
    """sig: [B, M, ld] float32 rows (zero beyond n1 / n2 for even / odd rows when they differ)."""
    b, m, ld = sig.shape
    p = len(pairs)
    n = n1 + n2 - 1
    k = np.full((b, p, num_peaks), -7, np.int32)
    cnt = np.zeros((b, p), np.int32)
    pk = np.zeros((b, p), np.float32)
    gm = np.zeros((b, p), np.float32)
    fl = np.zeros((b, p), np.uint32)
    corr = np.zeros((b, p, n), np.float32)
    lib().emu_generic_gcc_phat(int(use_double), _p(sig, C.c_float), C.c_longlong(b), m, ld, n1, n2, _p(pairs, C.c_int), p,
                               win_half, dist, method, C.c_float(mult), num_peaks, C.c_float(eps), _p(k, C.c_int),
                               _p(cnt, C.c_int), _p(pk, C.c_float), _p(gm, C.c_float), _p(fl, C.c_uint),
                               _p(corr, C.c_float))
    return k, cnt, pk, gm, fl, corr


def fft2_gcc_phat(sig, n1, n2, pairs, win_half, dist, plan_id=-1, method=0, mult=1.0, num_peaks=1, eps=0.0, fast=False,
                  smem_conv=False):
    """Same as generic_gcc_phat(use_double=False) on the second-generation convolution engine (pal_fft2.cuh);
    plan_id -1 picks the product's plan for n.  Returns the plan used as the last element."""
    b, m, ld = sig.shape
    p = len(pairs)
    n = n1 + n2 - 1
    k = np.full((b, p, num_peaks), -7, np.int32)
    cnt = np.zeros((b, p), np.int32)
    pk = np.zeros((b, p), np.float32)
    gm = np.zeros((b, p), np.float32)
    fl = np.zeros((b, p), np.uint32)
    corr = np.zeros((b, p, n), np.float32)
    used = lib().emu_fft2_gcc_phat(int(plan_id) + (200 if smem_conv else (100 if fast else 0)), _p(sig, C.c_float), C.c_longlong(b), m, ld, n1, n2, _p(pairs, C.c_int), p,
                                   win_half, dist, method, C.c_float(mult), num_peaks, C.c_float(eps), _p(k, C.c_int),
                                   _p(cnt, C.c_int), _p(pk, C.c_float), _p(gm, C.c_float), _p(fl, C.c_uint),
                                   _p(corr, C.c_float))
    if used < 0:
        raise ValueError(f"emu_fft2_gcc_phat: plan {plan_id} cannot hold n = {n} (rc {used})")
    return k, cnt, pk, gm, fl, corr, used


def _pd(a):
    return a.ctypes.data_as(C.POINTER(C.c_double)) if a is not None else None


def image_sources(sources, planes, plane_mat, mat_abs, mat_freq, mics, max_order, frequency, threshold, k_max,
                  round_scale=1e6):
    sources = np.ascontiguousarray(sources, np.float64).reshape(-1, 3)
    planes = np.ascontiguousarray(planes, np.float64).reshape(-1, 4)
    mics = np.ascontiguousarray(mics, np.float64).reshape(-1, 3)
    plane_mat = np.ascontiguousarray(plane_mat, np.int32)
    mat_abs = np.ascontiguousarray(mat_abs, np.float64)
    mat_freq = np.ascontiguousarray(mat_freq, np.float64)
    b = len(sources)
    pos = np.zeros((b, k_max, 3), np.float64)
    mat = np.zeros((b, k_max), np.int32)
    cnt = np.zeros(b, np.int32)
    lib().emu_image_sources(_pd(sources), C.c_longlong(b), _pd(planes), _p(plane_mat, C.c_int), len(planes), _pd(mat_abs),
                            _pd(mat_freq), _pd(mics), len(mics), max_order, C.c_double(frequency), C.c_double(threshold),
                            C.c_double(round_scale), k_max, _pd(pos), _p(mat, C.c_int), _p(cnt, C.c_int))
    return pos, mat, cnt


def path_table(src, img_pos, img_mat, mics, mat_abs, mat_freq, air_mat, frequency, c_sound):
    src = np.ascontiguousarray(src, np.float64)
    img_pos = np.ascontiguousarray(img_pos, np.float64).reshape(-1, 3)
    img_mat = np.ascontiguousarray(img_mat, np.int32)
    mics = np.ascontiguousarray(mics, np.float64).reshape(-1, 3)
    k1 = len(img_pos) + 1
    tau = np.zeros((len(mics), k1))
    gain = np.zeros((len(mics), k1))
    lib().emu_path_table(_pd(src), _pd(img_pos if len(img_pos) else np.zeros((1, 3))), _p(img_mat if len(img_mat) else np.zeros(1, np.int32), C.c_int),
                         len(img_pos), _pd(mics), len(mics), _pd(np.ascontiguousarray(mat_abs, np.float64)),
                         _pd(np.ascontiguousarray(mat_freq, np.float64)), air_mat, C.c_double(frequency),
                         C.c_double(c_sound), _pd(tau), _pd(gain))
    return tau, gain


def render_scene(base, total, tau, gain, fs, n_keep):
    base = np.ascontiguousarray(base, np.float32)
    tau = np.ascontiguousarray(tau, np.float64)
    gain = np.ascontiguousarray(gain, np.float64)
    m, k1 = tau.shape
    out = np.zeros((m, n_keep), np.float32)
    lib().emu_render_scene(_p(base, C.c_float), len(base), total, _pd(tau), _pd(gain), m, k1, C.c_double(fs), n_keep,
                           _p(out, C.c_float))
    return out


def path_table_batched(sources, img_pos, img_mat, img_count, mics, mat_abs, mat_freq, air_mat, frequency, c_sound):
    """sources [S,3], img_pos [S,K,3], img_mat [S,K], img_count [S]; mics [M,3] shared."""
    sources = np.ascontiguousarray(sources, np.float64).reshape(-1, 3)
    img_pos = np.ascontiguousarray(img_pos, np.float64)
    img_mat = np.ascontiguousarray(img_mat, np.int32)
    img_count = np.ascontiguousarray(img_count, np.int32)
    mics = np.ascontiguousarray(mics, np.float64).reshape(-1, 3)
    s, k = img_mat.shape
    ks = k + 1
    tau = np.zeros((s, len(mics), ks))
    gain = np.zeros((s, len(mics), ks))
    cnt = np.zeros(s, np.int32)
    mx = np.zeros(s)
    lib().emu_path_table_batched(_pd(sources), _pd(img_pos), _p(img_mat, C.c_int), _p(img_count, C.c_int), C.c_longlong(s), k,
                                 _pd(mics), len(mics), C.c_longlong(0), _pd(np.ascontiguousarray(mat_abs, np.float64)),
                                 _pd(np.ascontiguousarray(mat_freq, np.float64)), air_mat, C.c_double(frequency),
                                 C.c_double(c_sound), ks, _pd(tau), _pd(gain), _p(cnt, C.c_int), _pd(mx))
    return tau, gain, cnt, mx


def filtfilt_f64(x, b, a, zi, padlen):
    x = np.ascontiguousarray(x, np.float64)
    rows, n = x.shape
    nt = max(len(a), len(b))
    bb = np.zeros(nt); bb[:len(b)] = b
    aa = np.zeros(nt); aa[:len(a)] = a
    y = np.zeros_like(x)
    lib().emu_filtfilt_f64(_pd(x), C.c_longlong(rows), n, _pd(bb), _pd(aa), _pd(np.ascontiguousarray(zi, np.float64)), nt, padlen,
                           _pd(y))
    return y


def sync_align(sig, lens=None):
    """sig [S, M, ld] float64 -> ref_idx [S], peak_index [S, M], absmax [S, M], win [S, M, 5], energy [S, M]."""
    sig = np.ascontiguousarray(sig, np.float64)
    s, m, ld = sig.shape
    ln = None if lens is None else np.ascontiguousarray(lens, np.int32)
    ref = np.zeros(s, np.int32)
    pk = np.zeros((s, m), np.int32)
    am = np.zeros((s, m))
    win = np.zeros((s, m, 5))
    en = np.zeros((s, m))
    lib().emu_sync_align(_pd(sig), C.c_longlong(s), m, ld, _p(ln, C.c_int), _p(ref, C.c_int), _p(pk, C.c_int), _pd(am),
                         _pd(win), _pd(en))
    return ref, pk, am, win, en


def pad_rows_f64(x, pad, ld_out, lens=None):
    x = np.ascontiguousarray(x, np.float64)
    rows, ld = x.shape
    pad = np.ascontiguousarray(pad, np.int32)
    ln = None if lens is None else np.ascontiguousarray(lens, np.int32)
    out = np.full((rows, ld_out), np.nan)
    lib().emu_pad_rows_f64(_pd(x), C.c_longlong(rows), C.c_longlong(ld), _p(ln, C.c_int), _p(pad, C.c_int), _pd(out),
                           C.c_longlong(ld_out))
    return out


def solve_positions(mics, pairs, tdoa, c, weights=None, x0=None, lo=None, hi=None, buffer=5.0, max_iter=100, tol=1e-12):
    mics = np.ascontiguousarray(mics, np.float64)
    pairs = np.ascontiguousarray(pairs, np.int32)
    tdoa = np.ascontiguousarray(tdoa, np.float64)
    s_n, p_n = tdoa.shape
    m = mics.shape[-2]
    opt = lambda a: None if a is None else np.ascontiguousarray(a, np.float64)      # noqa: E731
    weights, x0, lo, hi = opt(weights), opt(x0), opt(lo), opt(hi)
    pos = np.zeros((s_n, 3))
    cost = np.zeros(s_n)
    it = np.zeros(s_n, np.int32)
    lib().emu_solve_positions(_pd(mics), C.c_longlong(3 * m if mics.ndim == 3 else 0), m, _p(pairs, C.c_int), p_n, _pd(tdoa),
                              _pd(weights), _pd(x0), _pd(lo), _pd(hi), C.c_longlong(s_n), C.c_double(c), C.c_double(buffer),
                              max_iter, C.c_double(tol), C.c_double(tol), C.c_double(tol), _pd(pos), _pd(cost), _p(it, C.c_int))
    return pos, cost, it
