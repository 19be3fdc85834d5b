"""CPU oracle for the PyAudioLocalization hot path (TEST INFRASTRUCTURE ONLY).

This module is a float64 numpy/scipy restatement of the reference's hot path
(stage 1: image-source multipath synthesis, stage 2: GCC-PHAT + bounded TDOA
pick).  It exists only so that `tests/`, `__graft_entry__.smoke()` and the
`cpu_baseline` / `--impl reference` legs of `bench.py` can check and time the
CUDA path against the reference algorithm.  Nothing in the product package
(`pyaudiolocalization_b200/`) may import it.

Parity pin: the reference ships no tests, fixtures or golden vectors
(SURVEY.md §4, §8c), so this oracle is pinned against OUTPUTS OF THE UNMODIFIED
REFERENCE generated in the build container by `tests/golden/make_golden.py`
(committed as `tests/golden/*.npz`) and re-checked in `tests/test_oracle.py`.

Two layers live here:

* "port" functions (`phat_correlation`, `get_time_delays_phat`,
  `fractional_delay`, `simulate_signals_with_multipath`, ...) perform the same
  numpy/scipy calls as the reference, in the same order, so they are
  bit-identical with it and cost the same CPU time (they are what the CPU
  baseline times);
* "restated" functions (`local_maxima_restated`, `tdoa_pick_restated`,
  `render_rows_restated`, ...) spell out the arithmetic hidden inside the
  third-party calls (scipy.signal.find_peaks 1.18, numpy.fft) in the form the
  CUDA kernels implement; tests prove them equal to the port layer.

Every function cites the reference file:line it follows (paths relative to
the reference repository root).
"""
from __future__ import annotations

import logging
import math
from typing import Any, Dict, List, Optional, Sequence, Tuple

import numpy as np

__all__ = [
    "DEFAULT_MATERIALS", "speed_of_sound", "reflect_point_across_plane", "distance",
    "calculate_attenuation", "generate_image_sources_iterative", "generate_signal",
    "fractional_delay", "normalize_signal", "dynamic_range_compression",
    "simulate_signals_with_multipath", "phat_correlation", "get_time_delays_phat",
    "pair_loop", "local_maxima_restated", "select_by_distance_restated",
    "window_half_width", "peak_distance", "tdoa_pick_restated", "tdoa_from_index",
    "path_table_restated", "render_rows_restated", "synchronize_signals_improved",
]

# materials.py:2-16 (data table, copied values)
DEFAULT_MATERIALS: Dict[str, Dict[str, float]] = {
    "air": {"absorption": 0.01, "freq": 0.1},
    "wood": {"absorption": 0.05, "freq": 0.8},
    "metal": {"absorption": 0.1, "freq": 0.6},
}


# --------------------------------------------------------------------------
# geometry / acoustics  (utils.py:15-106)
# --------------------------------------------------------------------------
def speed_of_sound(temperature: float, humidity: float, pressure: float = 101.325) -> float:
    """utils.py:15-27 — linear model with out-of-range clamps."""
    if not (-50 <= temperature <= 50):
        logging.warning("oracle: unusual temperature, using 20 C")
        temperature = 20
    if not (0 <= humidity <= 100):
        logging.warning("oracle: unusual humidity, using 50 %%")
        humidity = 50
    return 331 + 0.6 * temperature + 0.0124 * humidity + 0.0006 * (pressure - 101.325)


def reflect_point_across_plane(point: Sequence[float], plane: Sequence[float]) -> np.ndarray:
    """utils.py:29-42 — mirror image of `point` in plane a x + b y + c z + d = 0.

    Evaluation order is kept: ((a x + b y) + c z) + d, then 2*(...)/den, then
    x - a*factor (python scalar arithmetic, no FMA).
    """
    x, y, z = point
    a, b, c, d = plane
    den = a ** 2 + b ** 2 + c ** 2
    if den == 0:
        raise ValueError("invalid plane: a^2 + b^2 + c^2 == 0")
    f = 2 * (a * x + b * y + c * z + d) / den
    return np.array([x - a * f, y - b * f, z - c * f])


def distance(p1: Sequence[float], p2: Sequence[float]) -> float:
    """utils.py:44-48 — Euclidean norm of the difference."""
    return np.linalg.norm(np.array(p1) - np.array(p2))


def calculate_attenuation(d: float, material: str, frequency: float,
                          mats: Dict[str, Any]) -> float:
    """utils.py:50-65 — (1/d) * exp(-freq_factor*f*d) * exp(-absorption*d), d >= 0.1."""
    d = max(d, 0.1)
    if material not in mats:
        logging.warning("oracle: material %r undefined, using 'air'", material)
        material = "air"
    geo = 1 / d
    fa = np.exp(-mats[material]["freq"] * frequency * d)
    ab = np.exp(-mats[material]["absorption"] * d)
    return geo * fa * ab


def generate_image_sources_iterative(source, planes, max_order, frequency, mats,
                                     mic_positions, absorption_threshold=0.01,
                                     round_decimals=6) -> List[Dict[str, Any]]:
    """utils.py:67-106 — breadth-first image sources with de-dup and pruning.

    Discovery order, the 6-decimal de-dup key, the "failed candidates are not
    marked seen" rule and the last-plane material tag are all preserved.
    """
    out: List[Dict[str, Any]] = []
    frontier = [source]
    seen = {tuple(np.round(source, decimals=round_decimals))}
    for _order in range(1, max_order + 1):
        nxt = []
        for src in frontier:
            for pl in planes:
                img = reflect_point_across_plane(src, pl["plane"])
                key = tuple(np.round(img, decimals=round_decimals))
                if key in seen:
                    continue
                mat = pl.get("material", "air")
                if mat not in mats:
                    raise ValueError(f"material {mat!r} is not defined")
                if "absorption" not in mats[mat] or "freq" not in mats[mat]:
                    raise ValueError(f"material {mat!r} lacks absorption/freq")
                att = [calculate_attenuation(distance(img, m), mat, frequency, mats)
                       for m in mic_positions]
                if np.mean(att) > absorption_threshold and np.min(att) > absorption_threshold / 2:
                    seen.add(key)
                    out.append({"source": img, "material": mat})
                    nxt.append(img)
        frontier = nxt
        if not frontier:
            break
    return out


# --------------------------------------------------------------------------
# signal primitives  (signal_processing.py:25-36, 66-94)
# --------------------------------------------------------------------------
def generate_signal(signal_type: str, fs: float, duration: float, freq: float) -> np.ndarray:
    """signal_processing.py:25-36 — 'sine' | 'noise' | 'chirp' (deterministic ones only
    are used by parity tests; 'noise' draws from the global numpy RNG like the reference;
    'speech' is outside the parity scope because it is unseeded)."""
    t = np.linspace(0, duration, int(fs * duration), endpoint=False)
    if signal_type == "sine":
        return np.sin(2 * np.pi * freq * t)
    if signal_type == "noise":
        return np.random.normal(0, 1, size=t.shape)
    if signal_type == "chirp":
        from scipy.signal import chirp
        return chirp(t, f0=freq, f1=freq * 5, t1=duration, method="linear")
    raise ValueError("Unknown signal type. Available types: 'sine', 'noise', 'chirp', 'speech'")


def fractional_delay(signal: np.ndarray, delay: float, fs: float) -> np.ndarray:
    """signal_processing.py:66-80 — FFT(2N) linear-phase delay, real[:N], 1 % linear fades.

    The reference also builds a Hann window it never uses (:74); the call is kept so the
    port costs the same CPU time.
    """
    from scipy.signal import get_window
    n = len(signal)
    spec = np.fft.fft(signal, n=2 * n)
    f = np.fft.fftfreq(2 * n, d=1.0 / fs)
    y = np.fft.ifft(spec * np.exp(-1j * 2 * np.pi * f * delay)).real[:n]
    get_window("hann", n)
    fade = int(0.01 * n)
    w = np.ones(n)
    w[:fade] *= np.linspace(0, 1, fade)
    w[-fade:] *= np.linspace(1, 0, fade)   # like the reference: breaks when fade == 0 (n < 100)
    y *= w
    return y


def normalize_signal(x: np.ndarray) -> np.ndarray:
    """signal_processing.py:82-86."""
    m = np.max(np.abs(x))
    return x if m == 0 else x / m


def dynamic_range_compression(x: np.ndarray, threshold: float = 0.8, epsilon: float = 1e-8) -> np.ndarray:
    """signal_processing.py:88-94 — sign(x)·log1p(|x|/thr + eps), renormalised."""
    xn = normalize_signal(x)
    y = np.sign(xn) * np.log1p(np.abs(xn) / threshold + epsilon)
    m = np.max(np.abs(y))
    if m > 0:
        y /= m
    return y


def simulate_signals_with_multipath(source_pos, mic_positions, fs, c, duration=1.0,
                                    signal_type="sine", freq=1000, reflective_planes=None,
                                    material_properties=None, max_reflections=2,
                                    absorption_threshold=0.01, trim_to_duration=True,
                                    base_signal: Optional[np.ndarray] = None) -> List[np.ndarray]:
    """main.py:66-124 — per-mic sum of delayed, attenuated copies; trim; normalise; compress.

    `base_signal` (not in the reference signature) lets tests inject a seeded
    signal instead of the unseeded 'noise' generator.
    """
    base = generate_signal(signal_type, fs, duration, freq) if base_signal is None else base_signal
    images = generate_image_sources_iterative(source_pos, reflective_planes, max_reflections, freq,
                                              material_properties, mic_positions, absorption_threshold)
    max_delay = 0
    for mic in mic_positions:
        dmax = distance(source_pos, mic)
        for im in images:
            dmax = max(dmax, distance(im["source"], mic))
        max_delay = max(max_delay, dmax / c)
    total = int((duration + max_delay) * fs)
    padded = np.pad(base, (0, total - len(base)), "constant")
    out = []
    for mic in mic_positions:
        acc = np.zeros(total)
        d0 = distance(source_pos, mic)
        acc += fractional_delay(padded, d0 / c, fs) * calculate_attenuation(d0, "air", freq, material_properties)
        for im in images:
            d = distance(im["source"], mic)
            acc += fractional_delay(padded, d / c, fs) * calculate_attenuation(d, im["material"], freq,
                                                                               material_properties)
        if trim_to_duration:
            acc = acc[: int(duration * fs)]
        out.append(dynamic_range_compression(normalize_signal(acc)))
    return out


# --------------------------------------------------------------------------
# GCC-PHAT  (utils.py:108-181)
# --------------------------------------------------------------------------
def phat_correlation(sig1: np.ndarray, sig2: np.ndarray) -> np.ndarray:
    """utils.py:108-119 — ifft(F1·conj(F2)/(|F1·conj(F2)|+1e-10)).real, n = n1+n2-1, FFT order."""
    n = len(sig1) + len(sig2) - 1
    r = np.fft.fft(sig1, n=n) * np.conj(np.fft.fft(sig2, n=n))
    r /= np.abs(r) + 1e-10
    return np.fft.ifft(r).real


def get_time_delays_phat(sig1, sig2, fs, num_peaks=1, threshold_method="median",
                         threshold_multiplier=1.0, max_expected_delay=None
                         ) -> Tuple[List[float], np.ndarray, np.ndarray]:
    """utils.py:121-181 — PHAT correlation, scipy find_peaks(height, distance), window filter,
    fallbacks down to the unbounded global argmax."""
    from scipy.signal import correlation_lags, find_peaks
    corr = phat_correlation(sig1, sig2)
    t = correlation_lags(len(sig1), len(sig2), mode="full") / fs
    a = np.abs(corr)
    if threshold_method == "adaptive":
        thr = threshold_multiplier * (np.mean(a) + np.std(a))
    else:
        thr = threshold_multiplier * np.median(a)
    dist = int(fs * 0.001)
    peaks, props = find_peaks(corr, height=thr, distance=dist)
    if len(peaks) == 0:
        peaks, props = find_peaks(corr, height=np.mean(a), distance=dist)
        if len(peaks) == 0:
            return [t[np.argmax(corr)]], corr, t
    if max_expected_delay is not None:
        ok = [i for i in range(len(peaks)) if abs(t[peaks[i]]) <= max_expected_delay]
        if not ok:
            peaks, props = find_peaks(corr, height=np.mean(a), distance=dist)
            ok = [i for i in range(len(peaks)) if abs(t[peaks[i]]) <= max_expected_delay]
            if not ok:
                return [t[np.argmax(corr)]], corr, t
        peaks = peaks[ok]
        props["peak_heights"] = props["peak_heights"][ok]
    order = np.argsort(props["peak_heights"])[::-1]
    return list(t[peaks[order][:num_peaks]]), corr, t


def pair_loop(signals: Sequence[np.ndarray], fs: float, max_expected_delay=None,
              calib_delays: Optional[np.ndarray] = None):
    """main.py:195-231 — all i<j pairs with num_peaks=1; optional calibration correction;
    corr_matrix[i,j] = max(corr).  Returns (td_diffs, mic_pairs, corr_matrix)."""
    m = len(signals)
    tds, pairs = [], []
    cm = np.zeros((m, m))
    for i in range(m):
        for j in range(i + 1, m):
            td, corr, _ = get_time_delays_phat(signals[i], signals[j], fs, num_peaks=1,
                                               max_expected_delay=max_expected_delay)
            for v in td:
                if calib_delays is not None:
                    v = v - (calib_delays[j] - calib_delays[i])
                tds.append(v)
                pairs.append((i, j))
            cm[i, j] = cm[j, i] = np.max(corr)
    return tds, pairs, cm


# --------------------------------------------------------------------------
# restatements: what the kernels implement
# --------------------------------------------------------------------------
def local_maxima_restated(c: np.ndarray) -> np.ndarray:
    """scipy.signal._peak_finding_utils._local_maxima_1d semantics (called from
    find_peaks, utils.py:152): strict rise, optional plateau, strict fall; a plateau
    reports its floor-midpoint; the two end samples are never peaks."""
    n = len(c)
    out = []
    i = 1
    while i < n - 1:
        if c[i - 1] < c[i]:
            j = i + 1
            while j < n - 1 and c[j] == c[i]:
                j += 1
            if c[j] < c[i]:
                out.append((i + j - 1) // 2)
                i = j
        i += 1
    return np.asarray(out, dtype=np.intp)


def select_by_distance_restated(peaks: np.ndarray, heights: np.ndarray, dist: int) -> np.ndarray:
    """scipy _select_by_peak_distance: visit peaks highest-first (equal heights: the
    later index first, which is what reversing a stable ascending sort gives); a peak
    that is still alive deletes every other peak closer than `dist` samples."""
    order = np.argsort(heights, kind="stable")
    keep = np.ones(len(peaks), dtype=bool)
    for j in order[::-1]:
        if not keep[j]:
            continue
        k = j - 1
        while k >= 0 and peaks[j] - peaks[k] < dist:
            keep[k] = False
            k -= 1
        k = j + 1
        while k < len(peaks) and peaks[k] - peaks[j] < dist:
            keep[k] = False
            k += 1
    return keep


def peak_distance(fs: float) -> int:
    """utils.py:151 — int(fs * 0.001); scipy rejects values < 1 with ValueError."""
    d = int(fs * 0.001)
    if d < 1:
        raise ValueError("`distance` must be greater or equal to 1")
    return d


def window_half_width(n1: int, n2: int, fs: float, max_expected_delay: Optional[float]) -> int:
    """Largest |lag| (in samples) that passes `abs(lag/fs) <= max_expected_delay`
    (utils.py:163), decided with the same float64 division.  -1 = unbounded (None),
    -2 = empty window (negative or NaN bound: not even lag 0 passes)."""
    if max_expected_delay is None:
        return -1
    if not (0.0 <= max_expected_delay):
        return -2
    big = max(n1, n2)
    m = min(int(min(max_expected_delay * fs, float(big))) + 2, big)
    while m > 0 and not (abs(np.float64(m) / fs) <= max_expected_delay):
        m -= 1
    return m


def tdoa_pick_restated(c: np.ndarray, n2: int, win_half: int, dist: int, num_peaks: int = 1,
                       threshold_method: str = "median", threshold_multiplier: float = 1.0
                       ) -> List[int]:
    """utils.py:140-181 as an explicit index algorithm.

    Returns raw IFFT indices k (lag = k-(n2-1)).  Fact used: whether a peak survives the
    distance rule depends only on HIGHER peaks (which pass any threshold the peak itself
    passes), so survival can be decided once over all local maxima, independent of the
    height threshold in force.  The reference's control flow then reduces to:
      gpk = highest local maximum; thr_eff = thr if gpk >= thr, else mean|c| if
      gpk >= mean|c|, else -> first global argmax;  take surviving in-window peaks with
      height >= thr_eff; none -> retry with mean|c|; none -> first global argmax
      (unbounded); sort by height descending (ties: later index first).
    """
    a = np.abs(c)
    mean_abs = np.mean(a)
    if threshold_method == "adaptive":
        thr = threshold_multiplier * (mean_abs + np.std(a))
    else:
        thr = threshold_multiplier * np.median(a)
    garg = int(np.argmax(c))
    pk = local_maxima_restated(c)
    if len(pk) == 0:
        return [garg]
    h = c[pk]
    gpk = h.max()
    if gpk >= thr:
        thr_eff = thr
    elif gpk >= mean_abs:
        thr_eff = mean_abs
    else:
        return [garg]
    cand = select_by_distance_restated(pk, h, dist)
    if win_half != -1:
        cand &= np.abs(pk - (n2 - 1)) <= win_half      # win_half == -2 -> nothing passes
    sel = cand & (h >= thr_eff)
    if win_half != -1 and not sel.any():
        sel = cand & (h >= mean_abs)
        if not sel.any():
            return [garg]
    pk_s, h_s = pk[sel], h[sel]
    order = np.argsort(h_s, kind="stable")[::-1]
    return [int(k) for k in pk_s[order][:num_peaks]]


def tdoa_from_index(k: int, n2: int, fs: float) -> np.float64:
    """utils.py:141-142 — time_lags[k] = (k - (n2-1)) / fs as numpy computes it
    (int64 lag array true-divided by fs)."""
    return (np.arange(k - (n2 - 1), k - (n2 - 1) + 1) / fs)[0]


def path_table_restated(source_pos, images, mic_positions, fs, c, duration, freq, mats):
    """main.py:94-116 — per (mic, path) delay [s] and raw gain; N = int((dur+max_delay)*fs)."""
    paths = [(np.asarray(source_pos, float), "air")] + [(im["source"], im["material"]) for im in images]
    m = len(mic_positions)
    tau = np.zeros((m, len(paths)))
    gain = np.zeros((m, len(paths)))
    for i, mic in enumerate(mic_positions):
        for k, (p, mat) in enumerate(paths):
            d = distance(p, mic)
            tau[i, k] = d / c
            gain[i, k] = calculate_attenuation(d, mat, freq, mats)
    total = int((duration + tau.max()) * fs)
    return tau, gain, total


def render_rows_restated(base: np.ndarray, tau: np.ndarray, gain: np.ndarray, total: int,
                         fs: float, n_keep: int) -> np.ndarray:
    """The renderer in the form the GPU uses (SURVEY.md headline fact 6): every delayed
    copy shares one FFT and one fade window, so for mic i
        y_i = w · irfft( rfft(x_pad, 2N) · H_i )[:N],
        H_i[m] = Σ_k a_ik · exp(-j 2π m τ_ik fs / (2N)),  m = 0..N  (bin N uses f = -fs/2),
    followed by trim, max-normalise and the log compressor (main.py:119-122).  The
    per-mic gain is applied RELATIVE to the largest gain of that mic (raw gains sit at
    1e-38 with the stock materials; normalize_signal cancels any common factor)."""
    n = total
    x = np.zeros(n)
    x[: len(base)] = base
    spec = np.fft.rfft(x, 2 * n)
    mbin = np.arange(n + 1, dtype=np.float64)
    fade = int(0.01 * n)
    w = np.ones(n)
    w[:fade] *= np.linspace(0, 1, fade)
    w[-fade:] *= np.linspace(1, 0, fade)
    rows = []
    for i in range(tau.shape[0]):
        g = gain[i]
        gmax = g.max()
        rel = g / gmax if gmax > 0 else np.zeros_like(g)
        ph = mbin[:, None] * (tau[i][None, :] * fs / (2 * n))
        hresp = (rel[None, :] * np.exp(-2j * np.pi * ph)).sum(axis=1)
        # fftfreq puts bin N at -fs/2: phase factor exp(+j π fs τ); irfft keeps only its real part
        hresp[n] = (rel * np.cos(np.pi * fs * tau[i])).sum()
        y = np.fft.irfft(spec * hresp, 2 * n)[:n] * w
        y = y[:n_keep]
        rows.append(dynamic_range_compression(normalize_signal(y)))
    return np.asarray(rows)


# --------------------------------------------------------------------------
# between the stages: utils.py:407-457
# --------------------------------------------------------------------------
def synchronize_signals_improved(signals: Sequence[np.ndarray], fs: float,
                                 use_interpolation: bool = True) -> List[np.ndarray]:
    """Port of utils.synchronize_signals_improved (utils.py:407-457): same scipy calls in the same
    order.  Aligns every channel on the highest-energy one by the arg-max of the full
    cross-correlation, refined on a 5-point cubic spline, and left-pads."""
    from scipy.interpolate import CubicSpline
    from scipy.signal import correlate
    energies = [np.sum(sig ** 2) for sig in signals]                    # :415
    ref_idx = np.argmax(energies)
    reference = signals[ref_idx]
    ref_corr = correlate(reference, reference, mode='full')              # :418
    ref_peak = np.max(np.abs(ref_corr))
    shifts = []
    max_shift_samples = int(fs * 0.05)                                   # :421
    for idx, sig in enumerate(signals):
        if idx == ref_idx:
            shifts.append(0)
            continue
        corr = correlate(sig, reference, mode='full')                    # :426
        peak_index = np.argmax(np.abs(corr))
        if np.abs(corr[peak_index]) < 0.3 * ref_peak:                    # :428
            refined_peak = peak_index
        elif use_interpolation and peak_index > 1 and peak_index < len(corr) - 2:
            indices = np.arange(peak_index - 2, peak_index + 3)          # :433
            window_corr = corr[peak_index - 2: peak_index + 3]
            cs = CubicSpline(indices, window_corr)
            fine_indices = np.linspace(peak_index - 2, peak_index + 2, 100)
            fine_vals = cs(fine_indices)
            refined_peak = fine_indices[np.argmax(np.abs(fine_vals))]
        else:
            refined_peak = peak_index
        base_index = len(reference) - 1                                  # :441
        shift = refined_peak - base_index
        if abs(shift) > max_shift_samples:
            shift = 0
        shifts.append(shift)
    min_shift = min(shifts)                                              # :448
    adjusted = []
    for sig, shift in zip(signals, shifts):
        pad_left = max(0, int(round(shift - min_shift)))
        adjusted.append(np.pad(sig, (pad_left, 0), mode='constant'))
    max_length = max(len(s) for s in adjusted)
    return [np.pad(s, (0, max_length - len(s)), mode='constant') for s in adjusted]
