"""Drop-in mirror of the reference's `main` module API: `simulate_signals_with_multipath`
(main.py:66-124) and `localize_sound_source(config, ...)` (main.py:126-333).

The two data-parallel stages run on the GPU (scene synthesis: scene.py; GCC-PHAT / TDOA:
gcc_phat.py).  The two steps between them run on the GPU too: channel alignment (sync.py, utils.py:407-457) and
the Butterworth band-pass (filters.py, signal_processing.py:124-128).  What the reference keeps in
scipy/sklearn for the position solve -- clustering initialisation, bounded least squares and the
Differential-Evolution fallback -- stays on the host here as well (host_solver.py).
"""
from __future__ import annotations

import logging

import numpy as np
import torch

from . import gcc_phat as _g
from . import scene as _s
from . import sync as _sync
from .materials import material_properties  # noqa: F401
from .signal_processing import generate_signal
from .utils import speed_of_sound

config = {
    "fs": 44100,
    "duration": 1.0,
    "celsius": 20,
    "humidity": 50,
    "mic_positions": [[0.0, 0.0, 0.0], [1.0, 0.0, 0.0], [0.0, 1.0, 0.0], [0.0, 0.0, 1.0]],
    "source_position": [0.5, 0.5, 0.5],
    "signal_type": "sine",
    "freq": 1000,
    "reflective_planes": [
        {'plane': [1, 0, 0, -5], 'material': 'wood'},
        {'plane': [0, 1, 0, -5], 'material': 'metal'},
        {'plane': [0, 0, 1, -5], 'material': 'wood'},
    ],
    "calibration": {"signal_type": "chirp", "freq_start": 500, "freq_end": 5000, "attenuation_factor": 1.0,
                    "noise_level": 0.01},
    "localization": {"max_reflections": 3, "filter_method": "butterworth", "absorption_threshold": 0.01,
                     "analyze_correlation": True, "visualize_correlation": True, "clustering_method": "kmeans",
                     "clustering_eps": 0.001, "clustering_min_samples": 2, "max_expected_delay": 0.05},
}


def simulate_signals_device(source_pos, mic_positions, fs, c, duration=1.0, signal_type='sine', freq=1000,
                            reflective_planes=None, material_properties=None, max_reflections=2,
                            absorption_threshold=0.01, trim_to_duration=True, base_signal=None,
                            plan_cache=None) -> torch.Tensor:
    """Same computation as simulate_signals_with_multipath, result left on the device as a
    [M, n] float32 tensor.  `base_signal` optionally replaces generate_signal (e.g. a seeded
    noise burst)."""
    base = generate_signal(signal_type, fs, duration, freq) if base_signal is None else base_signal
    mats = material_properties
    planes = list(reflective_planes or [])
    n_img = 0
    k_max = None
    pos = mat = None
    table = _s.MaterialTable(mats, _s._dev())
    if max_reflections >= 1 and planes:
        while True:
            pos, mat, cnt, table = _s.image_sources_batched([source_pos], planes, max_reflections, freq, mats,
                                                            mic_positions, absorption_threshold, 6, k_max)
            n_img = int(cnt[0].item())
            if n_img >= 0:
                break
            k_max = pos.shape[1] * 4
        pos, mat = pos[0], mat[0]
    return _s.render_scene(base, source_pos, pos, mat, n_img, mic_positions, fs, c, duration, freq, table,
                           trim_to_duration=trim_to_duration)


def prepare_scenes_batched(source_positions, mic_positions, fs, c, duration=1.0, signal_type='sine', freq=1000,
                           reflective_planes=None, material_properties=None, max_reflections=2,
                           absorption_threshold=0.01, trim_to_duration=True, base_signal=None):
    """Geometry half of simulate_scenes_batched (image sources, path tables, transform lengths; ends with the
    renderer's one read-back).  Returns a scene.RenderJob for `scene.execute_render`; everything is issued on the
    current stream, so a sweep can prepare its next batch on a side stream while the GPU renders the current one."""
    base = generate_signal(signal_type, fs, duration, freq) if base_signal is None else base_signal
    mats = material_properties
    planes = list(reflective_planes or [])
    srcs = np.asarray(source_positions, dtype=np.float64).reshape(-1, 3)
    dev = _s._dev()
    if max_reflections >= 1 and planes:
        k_max = None
        while True:
            pos, mat, cnt, table = _s.image_sources_batched(srcs, planes, max_reflections, freq, mats, mic_positions,
                                                            absorption_threshold, 6, k_max)
            try:
                return _s.prepare_render(base, srcs, pos, mat, cnt, mic_positions, fs, c, duration, freq, table,
                                         trim_to_duration=trim_to_duration)
            except RuntimeError as e:          # image list overflow: retry with a larger k_max
                if "overflowed k_max" not in str(e):
                    raise
                k_max = pos.shape[1] * 4
    table = _s.MaterialTable(mats, dev)
    pos = torch.zeros((len(srcs), 1, 3), dtype=torch.float64, device=dev)
    mat = torch.zeros((len(srcs), 1), dtype=torch.int32, device=dev)
    cnt = torch.zeros((len(srcs),), dtype=torch.int32, device=dev)
    return _s.prepare_render(base, srcs, pos, mat, cnt, mic_positions, fs, c, duration, freq, table,
                             trim_to_duration=trim_to_duration)


def simulate_scenes_batched(source_positions, mic_positions, fs, c, duration=1.0, signal_type='sine', freq=1000,
                            reflective_planes=None, material_properties=None, max_reflections=2,
                            absorption_threshold=0.01, trim_to_duration=True, base_signal=None,
                            plan_cache=None) -> torch.Tensor:
    """simulate_signals_with_multipath (main.py:66-124) for MANY source positions / scenes in one go:
    source_positions [S, 3]; mic_positions [M, 3] (shared array) or [S, M, 3]; reflective_planes one
    list (shared room) or a list of S lists (one room per scene, same plane count and materials).
    Returns a [S, M, n] float32 tensor on the device.  Image sources, path tables, transfer
    functions, inverse transforms and the normalise / compress epilogue all run batched.  A sweep that calls
    this batch after batch passes one `scene.RenderPlanCache()` (and a device-resident `base_signal`) so that the
    per-length tables of the renderer are built once."""
    job = prepare_scenes_batched(source_positions, mic_positions, fs, c, duration, signal_type, freq, reflective_planes,
                                 material_properties, max_reflections, absorption_threshold, trim_to_duration, base_signal)
    return _s.execute_render(job, plan_cache=plan_cache)


def simulate_signals_with_multipath(source_pos, mic_positions, fs, c, duration=1.0, signal_type='sine', freq=1000,
                                    reflective_planes=None, material_properties=None, max_reflections=2,
                                    absorption_threshold=0.01, trim_to_duration=True):
    """main.py:66-124 -- list of M float64 arrays (one per microphone)."""
    out = simulate_signals_device(source_pos, mic_positions, fs, c, duration, signal_type, freq, reflective_planes,
                                  material_properties, max_reflections, absorption_threshold, trim_to_duration)
    host = out.double().cpu().numpy()
    return [host[i] for i in range(host.shape[0])]


def localize_sound_source(config, calibration_data=None, audio_files=None, use_simulation=True, show_plots=True):
    """main.py:126-333 -- same config keys, defaults, log lines, exceptions and result dict."""
    from . import host_solver as H
    fs = config["fs"]
    duration = config["duration"]
    mic_positions = np.array(config["mic_positions"])
    source_position = config["source_position"]
    signal_type = config["signal_type"]
    freq = config["freq"]
    reflective_planes = config.get("reflective_planes", [])
    material_props = material_properties
    lp = config.get("localization", {})
    filter_method = lp.get("filter_method", "butterworth")
    max_reflections = lp.get("max_reflections", 2)
    absorption_threshold = lp.get("absorption_threshold", 0.01)
    analyze_correlation = lp.get("analyze_correlation", False)
    visualize_correlation = lp.get("visualize_correlation", False)
    clustering_method = lp.get("clustering_method", "kmeans")
    clustering_eps = lp.get("clustering_eps", 0.001)
    clustering_min_samples = lp.get("clustering_min_samples", 2)
    max_expected_delay = lp.get("max_expected_delay", None)

    calib_delays = None
    if calibration_data is not None:
        if len(calibration_data) != len(mic_positions):
            logging.warning("Anzahl der Kalibrierdaten stimmt nicht mit der Anzahl der Mikrofone überein. Ignoriere Kalibrierung für diesen Durchlauf.")
        else:
            try:
                calib_delays = np.array([d.get('delay', 0.0) for d in calibration_data], dtype=float)
                logging.info("Kalibrierungskorrektur wird angewendet.")
            except Exception as e:          # noqa: BLE001 - same breadth as the reference
                logging.warning(f"Fehler beim Verarbeiten der Kalibrierdaten: {e}. Ignoriere Kalibrierung.")
                calib_delays = None

    c = speed_of_sound(config["celsius"], config["humidity"])
    logging.info(f"Berechnete Schallgeschwindigkeit: {c:.2f} m/s")

    if use_simulation:
        if source_position is None:
            raise ValueError("source_position muss angegeben werden, wenn use_simulation=True.")
        signals = simulate_signals_with_multipath(source_pos=source_position, mic_positions=mic_positions, fs=fs, c=c,
                                                  duration=duration, signal_type=signal_type, freq=freq,
                                                  reflective_planes=reflective_planes,
                                                  material_properties=material_props, max_reflections=max_reflections,
                                                  absorption_threshold=absorption_threshold, trim_to_duration=True)
        logging.info("Simulierte Signale erzeugt.")
    else:
        if audio_files is None:
            raise ValueError("Audio-Dateien müssen angegeben werden, wenn use_simulation=False.")
        if len(audio_files) != len(mic_positions):
            raise ValueError("Die Anzahl der Audio-Dateien muss mit der Anzahl der Mikrofone übereinstimmen.")
        signals = H.read_audio_files(audio_files, fs)
        logging.info("Echte Audiodaten geladen.")

    signals = _sync.synchronize_signals_improved(signals, fs)      # utils.py:407-457 on the GPU (sync.py)
    logging.info("Signale synchronisiert.")
    if filter_method == 'butterworth' and len({len(sig) for sig in signals}) == 1:
        # all channels through pal_filtfilt in ONE call (float64, bit-identical to scipy's filtfilt per channel)
        from .filters import noise_reduction_batched
        x = torch.as_tensor(np.ascontiguousarray(np.stack([np.asarray(sig, dtype=np.float64) for sig in signals]))).to(_s._dev())
        filtered = list(noise_reduction_batched(x, fs).cpu().numpy())
    else:
        filtered = [H.noise_reduction(sig, fs, method=filter_method) for sig in signals]
    for i in range(len(filtered)):
        logging.info(f"Signal {i+1} gefiltert mit '{filter_method}' Noise Reduction.")

    # ---- stage 2 on the GPU: every pair in one call (main.py:202-228) ------------------------
    m = len(filtered)
    n = len(filtered[0])
    # float64 rows -> float64 ingest (pal_gcc_phat_tdoa_f64): the filtered channels are band-limited, and PHAT gives the
    # stop-band bins unit weight -- rounding them to float32 first would change the correlation, not just its last digits
    frames = torch.from_numpy(np.ascontiguousarray(np.stack(filtered).astype(np.float64))[None]).to(_s._dev())
    res = _g.gcc_phat_tdoa_batched(frames, fs, max_expected_delay, num_peaks=1,
                                   return_corr=bool(analyze_correlation or visualize_correlation))
    td_all = res.tdoa_seconds()[0, :, 0]
    gmax = res.gmax[0].double().cpu().numpy()
    corr_rows = res.corr[0].double().cpu().numpy() if res.corr is not None else None
    pairs = _g.all_pairs(m)
    td_diffs, mic_pairs = [], []
    corr_matrix = np.zeros((m, m))
    correlation_metrics = {}
    corr_data_for_3d, pairs_for_3d = [], []
    for p, (i, j) in enumerate(pairs):
        i, j = int(i), int(j)
        td = td_all[p]
        if calib_delays is not None:
            correction = calib_delays[j] - calib_delays[i]
            td_corrected = td - correction
            td_diffs.append(td_corrected)
            mic_pairs.append((i, j))
            logging.info(f"Mikrofonpaar {i+1}-{j+1}: TDOA gemessen={td:.6f}s, Korrektur={correction:+.6f}s, TDOA korrigiert={td_corrected:.6f}s")
        else:
            td_diffs.append(td)
            mic_pairs.append((i, j))
            logging.info(f"Zeitdifferenz für Mikrofonpaar {i+1}-{j+1}: {td:.6f} s (ohne Kalibrierung)")
        if analyze_correlation:
            metrics = H.compute_cross_correlation_metrics(corr_rows[p], filtered[i], filtered[j], fs, alpha=0.05)
            correlation_metrics[(i, j)] = metrics
            logging.info(f"Cross-Correlation-Metriken für Mikrofonpaar {i+1}-{j+1}: {metrics}")
        corr_matrix[i, j] = corr_matrix[j, i] = gmax[p]
        if visualize_correlation:
            corr_data_for_3d.append(corr_rows[p])
            pairs_for_3d.append((i, j))
    if not mic_pairs:
        raise RuntimeError("Keine gültigen Mikrofonpaare mit ermittelten Zeitverzögerungen.")
    del n

    x_source, y_source, z_source = H.solve_position(mic_positions, mic_pairs, td_diffs, c, correlation_metrics,
                                                    analyze_correlation, clustering_method, clustering_eps,
                                                    clustering_min_samples)
    H.maybe_plot(use_simulation, visualize_correlation, show_plots, mic_positions, source_position,
                 (x_source, y_source, z_source), corr_matrix, corr_data_for_3d, pairs_for_3d, fs)
    if analyze_correlation:
        logging.info("Erweiterte Cross-Correlation Metriken:")
        for pair, metrics in correlation_metrics.items():
            logging.info(f"Mikrofonpaar {pair[0]+1}-{pair[1]+1}: {metrics}")
    return {
        "estimated_position": np.array([x_source, y_source, z_source]),
        "actual_position": source_position if use_simulation else None,
        "mic_positions": mic_positions,
        "correlation_metrics": correlation_metrics if analyze_correlation else None,
        "correlation_matrix": corr_matrix if visualize_correlation else None,
        "calibration_data": calibration_data,
    }
