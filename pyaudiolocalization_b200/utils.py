"""Drop-in mirror of the reference's `utils` functions on the hot path (same names, arguments,
return types and error behaviour), computed on the GPU through libpal_b200.so.

Reference: utils.py:15-181.  Scalar geometry helpers (speed_of_sound, reflect_point_across_plane,
distance, calculate_attenuation) are trivial host arithmetic in the reference and stay host
arithmetic here; everything that costs time (image sources, rendering, GCC-PHAT, peak picking)
goes to the device.  There is no CPU fallback for those.
"""
from __future__ import annotations

import logging
from typing import Any, Dict, List, Optional, Tuple

import numpy as np
import torch

from . import gcc_phat as _g
from .materials import material_properties  # noqa: F401  (re-exported like the reference)


def speed_of_sound(temperature: float, humidity: float, pressure: float = 101.325) -> float:
    """utils.py:15-27."""
    if temperature < -50 or temperature > 50:
        logging.warning("Ungewöhnliche Temperatur. Verwende Standardwert 20°C.")
        temperature = 20
    if humidity < 0 or humidity > 100:
        logging.warning("Ungewöhnliche Luftfeuchtigkeit. Verwende Standardwert 50%.")
        humidity = 50
    return 331 + 0.6 * temperature + 0.0124 * humidity + 0.0006 * (pressure - 101.325)


def reflect_point_across_plane(point, plane) -> np.ndarray:
    """utils.py:29-42 (ValueError for a zero normal)."""
    x, y, z = point
    a, b, c, d = plane
    den = a ** 2 + b ** 2 + c ** 2
    if den == 0:
        raise ValueError("Ungültige Ebene: a^2 + b^2 + c^2 ist 0.")
    f = 2 * (a * x + b * y + c * z + d) / den
    return np.array([x - a * f, y - b * f, z - c * f])


def distance(point1, point2) -> float:
    """utils.py:44-48."""
    return np.linalg.norm(np.array(point1) - np.array(point2))


def calculate_attenuation(distance_val: float, material: str, frequency: float,
                          material_properties: Dict[str, Any]) -> float:
    """utils.py:50-65 (unknown material: warning, then 'air')."""
    distance_val = max(distance_val, 0.1)
    if material not in material_properties:
        logging.warning(f"Material '{material}' nicht definiert. Nutze 'air' als Standard.")
        material = 'air'
    m = material_properties[material]
    return (1 / distance_val) * np.exp(-m['freq'] * frequency * distance_val) * np.exp(-m['absorption'] * distance_val)


def _device():
    if not torch.cuda.is_available():
        raise RuntimeError("pyaudiolocalization_b200 needs a CUDA device; there is no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


def _pair_rows(sig1, sig2):
    """The two signals as one [1, 2, max(n1, n2)] device tensor.  float32 inputs stay float32 (the batched kernels'
    type); anything else is taken as float64, the reference's arithmetic type, and goes through the float64 ingest
    (pal_gcc_phat_tdoa_f64): the single-call drop-in functions then decide exactly like the reference."""
    a, b = np.asarray(sig1), np.asarray(sig2)
    dt = np.float32 if (a.dtype == np.float32 and b.dtype == np.float32) else np.float64
    a = np.ascontiguousarray(a.astype(dt, copy=False).reshape(-1))
    b = np.ascontiguousarray(b.astype(dt, copy=False).reshape(-1))
    n1, n2 = len(a), len(b)
    if n1 < 1 or n2 < 1:
        raise ValueError("empty signal")
    rows = np.zeros((1, 2, max(n1, n2)), dt)
    rows[0, 0, :n1] = a
    rows[0, 1, :n2] = b
    return torch.from_numpy(rows).to(_device()), n1, n2


def phat_correlation(sig1: np.ndarray, sig2: np.ndarray) -> np.ndarray:
    """utils.py:108-119 -- PHAT cross-correlation, length n1+n2-1, FFT order, float64 array (float64 inputs are
    transformed in float64 on the device and returned through a float32 buffer: within 1e-6 of the reference)."""
    rows, n1, n2 = _pair_rows(sig1, sig2)
    res = _g.gcc_phat_tdoa_batched(rows, 16000.0, None, return_corr=True, refine=False,
                                   lengths=(n1, n2))
    return res.corr[0, 0].double().cpu().numpy()


def get_time_delays_phat(sig1: np.ndarray, sig2: np.ndarray, fs: float, num_peaks: int = 1,
                         threshold_method: str = 'median', threshold_multiplier: float = 1.0,
                         max_expected_delay: Optional[float] = None
                         ) -> Tuple[List[float], np.ndarray, np.ndarray]:
    """utils.py:121-181 -- (time delays, correlation, lags in seconds).  The delays are
    (k - (n2-1)) / fs evaluated in float64 from the integer peak indices, exactly as the
    reference's `time_lags[selected_peaks]`."""
    rows, n1, n2 = _pair_rows(sig1, sig2)
    res = _g.gcc_phat_tdoa_batched(rows, fs, max_expected_delay, num_peaks=num_peaks,
                                   threshold_method=threshold_method,
                                   threshold_multiplier=threshold_multiplier, return_corr=True,
                                   lengths=(n1, n2))
    cnt = int(res.k_count[0, 0].item())
    k = res.k_idx[0, 0, :cnt].cpu().numpy().astype(np.int64)
    flags = int(res.flags[0, 0].item())
    if flags & _g._lib.FLAG_ALT_THRESHOLD:
        logging.warning(f"Keine Peaks mit Schwellenwertmethode '{threshold_method}' gefunden. Versuche alternativen Schwellenwert.")
    if flags & _g._lib.FLAG_FALLBACK_ARGMAX:
        logging.warning("Keine gültigen Peaks. Nutze Maximum der Korrelation als Verzögerung.")
    time_lags = np.arange(-(n2 - 1), n1) / fs          # scipy.signal.correlation_lags(n1, n2, 'full') / fs
    corr = res.corr[0, 0].double().cpu().numpy()
    return list(time_lags[k]), corr, time_lags


def generate_image_sources_iterative(source, planes, max_order: int, frequency: float,
                                     material_properties: Dict[str, Any], mic_positions,
                                     absorption_threshold: float = 0.01, round_decimals: int = 6
                                     ) -> List[Dict[str, Any]]:
    """utils.py:67-106 -- list of {'source': ndarray[3], 'material': str} in discovery order,
    computed by one thread block on the device (float64, the reference's evaluation order)."""
    from . import scene as _s
    if max_order < 1 or not planes:
        return []
    k_max = None
    while True:
        pos, mat, cnt, table = _s.image_sources_batched([source], planes, max_order, frequency, material_properties,
                                                        mic_positions, absorption_threshold, round_decimals, k_max)
        n = int(cnt[0].item())
        if n >= 0:
            break
        k_max = pos.shape[1] * 4            # more images than the first allocation: retry larger
    p = pos[0, :n].cpu().numpy()
    m = mat[0, :n].cpu().numpy()
    return [{'source': p[i].copy(), 'material': table.names[int(m[i])]} for i in range(n)]


# Host-side helpers of the reference's `utils` namespace (unchanged algorithms, see host_solver.py)
from .host_solver import (bootstrap_significance, compute_cross_correlation_metrics,  # noqa: E402,F401
                          compute_peak_to_peak_ratio, compute_snr, compute_weights,
                          determine_optimal_number_of_clusters, dynamic_bounds_extended, equations,
                          heuristic_initialization_adaptive, read_audio_files)
# utils.py:407-457 on the GPU (sync.py)
from .sync import synchronize_signals_improved  # noqa: E402,F401
