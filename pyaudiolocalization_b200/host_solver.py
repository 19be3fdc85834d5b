"""Host-side remainder of `localize_sound_source` (the parts the reference runs in scipy /
sklearn and that stay on the host by design: SURVEY.md section 2 "OUT OF SCOPE", north_star
"Differential Evolution / least-squares position solving stays on the host").

Same algorithms and defaults as the reference (utils.py:183-497, signal_processing.py:109-138,
main.py:238-319), written against numpy/scipy/sklearn.  One deliberate difference: the bootstrap
significance test (utils.py:183-216), which is `phat_correlation` 1000 times per pair and 99.8 %
of the README example's wall time, sends its 1000 permuted correlations to the GPU in one batch.
"""
from __future__ import annotations

import logging
import os
from typing import Any, Dict, List, Tuple

import numpy as np


# ------------------------------------------------------------------ pre-processing (main.py:185-191)
def read_audio_files(audio_files, expected_fs):
    """utils.py:459-482 (needs the optional `soundfile` / `resampy` packages)."""
    from .signal_processing import dynamic_range_compression, normalize_signal
    out = []
    for f in audio_files:
        if not os.path.isfile(f):
            logging.error(f"Audio file nicht gefunden: {f}")
            raise FileNotFoundError(f"Audio file nicht gefunden: {f}")
        try:
            import soundfile as sf
            sig, fs = sf.read(f)
            if sig.ndim > 1:
                sig = np.mean(sig, axis=1)
            if fs != expected_fs:
                import resampy
                logging.info(f"Resampling von '{f}' von {fs} Hz auf {expected_fs} Hz.")
                sig = resampy.resample(sig, fs, expected_fs, filter='kaiser_best')
            out.append(dynamic_range_compression(normalize_signal(sig)))
        except Exception as e:      # noqa: BLE001 - the reference wraps everything
            logging.error(f"Fehler beim Lesen der Audio-Datei '{f}': {e}")
            raise RuntimeError(f"Fehler beim Lesen der Audio-Datei '{f}': {e}")
    return out


def noise_reduction(signal, fs, method='butterworth', lowcut=300, highcut=3400, filter_order=101):
    """signal_processing.py:109-138."""
    from scipy.signal import butter, filtfilt, firwin, wiener
    nyq = 0.5 * fs
    if method == 'butterworth':
        # the Butterworth band-pass of main.py:191 runs on the device (pal_filtfilt): float64,
        # bit-identical to scipy.signal.filtfilt; the design call below is the reference's own
        import torch
        from .filters import filtfilt_batched
        b, a = butter(5, [lowcut / nyq, highcut / nyq], btype='band')
        x = torch.as_tensor(np.ascontiguousarray(np.asarray(signal, dtype=np.float64))).cuda()
        return filtfilt_batched(x, b, a).cpu().numpy()
    if method == 'fir':
        return filtfilt(firwin(filter_order, [lowcut / nyq, highcut / nyq], pass_zero=False), [1.0], signal)
    if method == 'wiener':
        return wiener(signal)
    raise ValueError("Unknown filter method. Available methods: 'butterworth', 'fir', 'wiener'")


# ------------------------------------------------------------------ correlation metrics (utils.py:183-271)
def bootstrap_significance(sig1, sig2, fs, num_bootstrap=1000, alpha=0.05, bootstrap_mode="permutation",
                           block_size=50):
    """utils.py:183-216 -- the resampling is drawn on the host from numpy's global RNG exactly
    like the reference; the `num_bootstrap` PHAT correlations and their maxima run as ONE batched
    GPU call (channel 0 = sig1, channels 1.. = resampled sig2)."""
    import torch

    from . import gcc_phat as g
    n1, n2 = len(sig1), len(sig2)
    rows = np.zeros((1, num_bootstrap + 1, max(n1, n2)), np.float32)
    rows[0, 0, :n1] = sig1
    for b in range(num_bootstrap):
        if bootstrap_mode == "permutation":
            s = np.random.permutation(sig2)
        elif bootstrap_mode == "block":
            nb = int(np.ceil(n2 / block_size))
            blocks = [sig2[i * block_size:(i + 1) * block_size] for i in range(nb)]
            np.random.shuffle(blocks)
            s = np.concatenate(blocks)[:n2]
        elif bootstrap_mode == "circular":
            s = np.roll(sig2, np.random.randint(0, n2))
        else:
            raise ValueError("Unbekannter bootstrap_mode. Nutze 'permutation', 'block' oder 'circular'.")
        rows[0, b + 1, :n2] = s
    if n1 != n2:
        # unequal lengths cannot share one batch layout; fall back to per-pair calls (still on the GPU)
        from .utils import phat_correlation
        peaks = [np.max(phat_correlation(sig1, rows[0, b + 1, :n2])) for b in range(num_bootstrap)]
        return np.percentile(peaks, 100 * (1 - alpha))
    pairs = np.array([(0, b + 1) for b in range(num_bootstrap)], np.int32)
    res = g.gcc_phat_tdoa_batched(torch.from_numpy(rows).cuda(), max(fs, 1000.0), None, pairs=pairs, refine=False)
    return np.percentile(res.gmax[0].double().cpu().numpy(), 100 * (1 - alpha))


def compute_peak_to_peak_ratio(corr):
    """utils.py:228-236."""
    trough = np.min(corr)
    return np.inf if trough == 0 else np.max(corr) / abs(trough)


def compute_snr(corr):
    """utils.py:238-251."""
    pk = int(np.argmax(corr))
    w = max(1, int(0.01 * len(corr)))
    noise = np.std(np.concatenate((corr[:max(0, pk - w)], corr[min(len(corr), pk + w):])))
    return np.inf if noise == 0 else np.max(corr) / noise


def compute_cross_correlation_metrics(corr, sig1, sig2, fs, alpha=0.05) -> Dict[str, Any]:
    """utils.py:261-271 (+ :218-226, :252-259): peak-to-peak ratio, SNR, bootstrap significance."""
    snr = compute_snr(corr)
    peak = np.max(corr)
    significant = bool(peak > bootstrap_significance(sig1, sig2, fs, alpha=alpha)) and snr > 2.0
    return {'peak_to_peak_ratio': compute_peak_to_peak_ratio(corr), 'snr': snr, 'significant': significant}


def compute_weights(correlation_metrics, mic_pairs):
    """utils.py:484-497."""
    w = np.array([(correlation_metrics.get(p) or {}).get('snr', 1.0) for p in mic_pairs])
    return w / np.mean(w) if np.mean(w) != 0 else w


# ------------------------------------------------------------------ position solve (main.py:238-298)
def determine_optimal_number_of_clusters(data, max_clusters=5, method='kmeans', eps=0.001, min_samples=2):
    """utils.py:273-302."""
    from sklearn.cluster import DBSCAN, KMeans
    from sklearn.metrics import silhouette_score
    x = np.array(data)
    if len(x) < 2:
        return 1
    if method == 'kmeans':
        best, best_k = -1, 1
        for k in range(2, min(max_clusters, len(x)) + 1):
            score = silhouette_score(x, KMeans(n_clusters=k, random_state=0).fit(x).labels_)
            if score > best:
                best, best_k = score, k
        return best_k
    if method == 'dbscan':
        labels = DBSCAN(eps=eps, min_samples=min_samples).fit(x).labels_
        ok = labels != -1
        if np.sum(ok) < 2:
            return 1
        return len(set(labels[ok])) if silhouette_score(x[ok], labels[ok]) > 0 else 1
    raise ValueError("Unbekannte Clustering-Methode. Verfügbare Methoden: 'kmeans', 'dbscan'")


def heuristic_initialization_adaptive(mic_positions, mic_pairs, tdoas, c, clustering_method='kmeans', eps=0.001,
                                      min_samples=2):
    """utils.py:304-362 -- hyperbola mid-points per pair, clustered into start positions."""
    from sklearn.cluster import DBSCAN, KMeans
    mp = np.array(mic_positions)
    centre = np.mean(mp, axis=0)
    if np.size(tdoas) == 0:
        return [centre.tolist()]
    est = []
    for (i, j), td in zip(mic_pairs, np.array(tdoas)):
        d = mp[j] - mp[i]
        nd = np.linalg.norm(d)
        if nd == 0:
            continue
        off = (c * abs(td)) / 2 * (d / nd)
        est.append(((mp[i] + mp[j]) / 2 + (-off if td > 0 else off)).tolist())
    if not est:
        return [centre.tolist()]
    if clustering_method == 'kmeans':
        k = determine_optimal_number_of_clusters(est, method='kmeans', eps=eps, min_samples=min_samples)
        guesses = KMeans(n_clusters=k, random_state=0).fit(est).cluster_centers_.tolist()
    elif clustering_method == 'dbscan':
        labels = DBSCAN(eps=eps, min_samples=min_samples).fit(est).labels_
        guesses = [np.mean([e for e, l in zip(est, labels) if l == lab], axis=0).tolist()
                   for lab in sorted(set(labels) - {-1})] or [centre.tolist()]
    else:
        guesses = [centre.tolist()]
    if not any(np.allclose(centre, g, atol=1e-6) for g in guesses):
        guesses.append(centre.tolist())
    return guesses


def dynamic_bounds_extended(mic_positions, tdoas, c, buffer=5.0):
    """utils.py:364-382."""
    mp = np.array(mic_positions)
    extra = max(np.percentile(c * np.abs(np.array(tdoas)), 75), 1.0) if np.size(tdoas) > 0 else 0.0
    lo, hi = np.min(mp, axis=0) - (buffer + extra), np.max(mp, axis=0) + (buffer + extra)
    return [(lo[i], hi[i]) for i in range(mp.shape[1] if mp.ndim > 1 else 1)]


def equations(vars, mic_positions, mic_pairs, tdoas, c, weights=None):
    """utils.py:384-405 -- residuals (d_j - d_i) - c * td, optionally weighted."""
    if weights is not None and len(weights) != len(mic_pairs):
        raise ValueError("Länge der Gewichte muss der Anzahl der Mikrofonpaare entsprechen.")
    src = np.array(vars)
    mp = np.asarray(mic_positions, dtype=float)
    out = []
    for idx, ((i, j), td) in enumerate(zip(mic_pairs, tdoas)):
        r = (np.linalg.norm(src - mp[j]) - np.linalg.norm(src - mp[i])) - c * td
        out.append(r * weights[idx] if weights is not None else r)
    return out


def solve_position(mic_positions, mic_pairs, td_diffs, c, correlation_metrics, analyze_correlation,
                   clustering_method, clustering_eps, clustering_min_samples) -> Tuple[float, float, float]:
    """main.py:233-298 -- heuristic starts, bounded trust-region least squares per start, and the
    Differential-Evolution fallback when every start fails."""
    from scipy.optimize import differential_evolution, least_squares
    for (i, j), td in zip(mic_pairs, td_diffs):
        logging.info(f"Differenz der Distanzen für Mikrofonpaar {i+1}-{j+1}: {c * td:.3f} m")
    guesses = heuristic_initialization_adaptive(mic_positions, mic_pairs, td_diffs, c, clustering_method=clustering_method,
                                                eps=clustering_eps, min_samples=clustering_min_samples)
    logging.info(f"Heuristisch initiale Positionen: {guesses}")
    bounds = dynamic_bounds_extended(mic_positions, td_diffs, c, buffer=5.0)
    lo, hi = [b[0] for b in bounds], [b[1] for b in bounds]
    guesses = [np.array([np.clip(g[k], lo[k], hi[k]) for k in range(len(g))]) for g in guesses]
    weights = compute_weights(correlation_metrics, mic_pairs) if (analyze_correlation and correlation_metrics) \
        else np.ones(len(mic_pairs))
    best, best_cost = None, np.inf
    for g in guesses:
        r = least_squares(equations, g, args=(mic_positions, mic_pairs, td_diffs, c, weights), bounds=(lo, hi),
                          method='trf', ftol=1e-6, xtol=1e-6, gtol=1e-6)
        if r.success and r.cost < best_cost:
            best, best_cost = r, r.cost
    if best is not None:
        x, y, z = best.x
        logging.info(f"Geschätzte Quelle: ({x:.3f}, {y:.3f}, {z:.3f}) m")
        return x, y, z
    logging.warning("Least Squares Optimierung fehlgeschlagen, versuche Differential Evolution.")
    r = differential_evolution(lambda v: np.sum(np.square(equations(v, mic_positions, mic_pairs, td_diffs, c, weights))),
                               bounds=list(zip(lo, hi)), strategy='best1bin', maxiter=1000, popsize=15, tol=1e-6,
                               mutation=(0.5, 1), recombination=0.7, polish=True, init='latinhypercube')
    if r.success:
        x, y, z = r.x
        logging.info(f"Geschätzte Quelle (Differential Evolution): ({x:.3f}, {y:.3f}, {z:.3f}) m")
        return x, y, z
    logging.error("Differential Evolution Optimierung fehlgeschlagen. Verwende den ersten initialen Schätzwert als Fallback.")
    x, y, z = guesses[0]
    return x, y, z


# ------------------------------------------------------------------ plots (main.py:300-319, plotting.py)
def maybe_plot(use_simulation, visualize_correlation, show_plots, mic_positions, source_position, estimate,
               corr_matrix, corr_rows, pairs, fs):
    """The reference's scatter / heat-map / 3-D plots.  matplotlib is an optional dependency:
    without it the figures are skipped with a log line (the reference would fail at import)."""
    if not (use_simulation or visualize_correlation):
        return
    try:
        import matplotlib
        if not show_plots:
            matplotlib.use("Agg")
        import matplotlib.pyplot as plt
    except Exception:      # noqa: BLE001
        logging.info("matplotlib nicht verfügbar: Plots werden übersprungen.")
        return
    if use_simulation:
        fig = plt.figure()
        ax = fig.add_subplot(111, projection='3d')
        ax.scatter(mic_positions[:, 0], mic_positions[:, 1], mic_positions[:, 2], c='r', marker='o', label='Mikrofone')
        ax.scatter(*source_position, c='g', marker='*', s=100, label='Tatsächliche Quelle')
        ax.scatter(*estimate, c='b', marker='x', s=100, label='Geschätzte Quelle')
        ax.set_xlabel('X (m)'); ax.set_ylabel('Y (m)'); ax.set_zlabel('Z (m)')
        ax.legend()
        plt.title('Sound Source Localization')
        plt.show() if show_plots else plt.savefig("localization_result.png")
        plt.close(fig)
    if visualize_correlation:
        fig = plt.figure()
        plt.imshow(corr_matrix, cmap='viridis', interpolation='nearest')
        plt.colorbar(label='Peak Correlation')
        plt.title('Correlation Heatmap')
        plt.show() if show_plots else plt.savefig("heatmap.png")
        plt.close(fig)
        fig = plt.figure()
        ax = fig.add_subplot(111, projection='3d')
        for idx, (row, pair) in enumerate(zip(corr_rows, pairs)):
            lags = np.arange(-(len(row) // 2), len(row) - len(row) // 2) / fs
            ax.plot(lags, np.full_like(lags, idx), row, label=f"{pair[0]+1}-{pair[1]+1}")
        ax.set_xlabel('Lag (s)'); ax.set_ylabel('Pair'); ax.set_zlabel('Correlation')
        plt.show() if show_plots else plt.savefig("correlation_3d.png")
        plt.close(fig)
