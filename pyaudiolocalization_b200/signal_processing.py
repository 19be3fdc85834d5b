"""Drop-in mirror of the reference's `signal_processing` functions on the hot path
(signal_processing.py:25-36, 66-94).  Signal GENERATION stays on the host like in the reference
(one base signal per scene, uploaded once); delays, normalisation and compression run on the GPU."""
from __future__ import annotations

import numpy as np
import torch

from . import scene as _s


def generate_signal(signal_type: str, fs: float, duration: float, freq: float) -> np.ndarray:
    """signal_processing.py:25-36 (host side; 'noise' and 'speech' draw from numpy's global RNG
    exactly like the reference)."""
    t = np.linspace(0, duration, int(fs * duration), endpoint=False)
    if signal_type == 'sine':
        return np.sin(2 * np.pi * freq * t)
    elif signal_type == 'noise':
        return np.random.normal(0, 1, size=t.shape)
    elif signal_type == 'chirp':
        from scipy.signal import chirp
        return chirp(t, f0=freq, f1=freq * 5, t1=duration, method='linear')
    elif signal_type == 'speech':
        return generate_realistic_speech(fs, duration)
    raise ValueError("Unknown signal type. Available types: 'sine', 'noise', 'chirp', 'speech'")


def generate_pink_noise(fs: float, duration: float) -> np.ndarray:
    """signal_processing.py:11-23 (host side, unseeded like the reference)."""
    n = int(fs * duration)
    spec = np.fft.rfft(np.random.randn(n))
    f = np.fft.rfftfreq(n, d=1. / fs)
    scale = np.ones_like(f)
    scale[1:] = 1 / np.sqrt(f[1:])
    scale[0] = 0
    pink = np.fft.irfft(spec * scale, n=n)
    return dynamic_range_compression(normalize_signal(pink))


def generate_realistic_speech(fs: float, duration: float) -> np.ndarray:
    """signal_processing.py:38-64 (host side, unseeded like the reference)."""
    from scipy.signal import get_window
    t = np.linspace(0, duration, int(fs * duration), endpoint=False)
    s = (1.0 * np.sin(2 * np.pi * 800 * t) + 0.8 * np.sin(2 * np.pi * 1150 * t + np.pi / 4)
         + 0.5 * np.sin(2 * np.pi * 2900 * t + np.pi / 2)) * get_window('hann', len(t))
    tr = np.zeros_like(t)
    ts = int(0.01 * fs)
    for _ in range(int(duration * 5)):
        i0 = np.random.randint(0, len(t) - ts)
        tr[i0:i0 + ts] += np.random.normal(0, 1, ts) * np.hanning(ts)
    s = s + tr + generate_pink_noise(fs, duration) * 0.05
    return dynamic_range_compression(normalize_signal(s))


def fractional_delay(signal: np.ndarray, delay: float, fs: float) -> np.ndarray:
    """signal_processing.py:66-80 -- FFT(2N) linear-phase delay, real[:N], 1 % linear fades,
    evaluated by the renderer kernels with a single unit-gain path."""
    x = np.asarray(signal, dtype=np.float64)
    n = len(x)
    if int(0.01 * n) < 1:
        raise ValueError("operands could not be broadcast together with shapes (0,) (%d,)" % n)
    dev = _s._dev()
    from . import _lib
    import ctypes as C
    L = _lib.lib()
    base = torch.from_numpy(x.astype(np.float32)).to(dev)
    tau = torch.tensor([[float(delay)]], dtype=torch.float64, device=dev)
    gain = torch.ones((1, 1), dtype=torch.float64, device=dev)
    out = torch.empty((1, n), dtype=torch.float32, device=dev)
    full, small = C.c_size_t(0), C.c_size_t(0)
    _lib.check(L.pal_render_workspace(n, 1, C.byref(full), C.byref(small)), "pal_render_workspace")
    ws, wp, wl = _s._ws(full.value, dev)
    _lib.check(L.pal_render_scene(base.data_ptr(), n, n, tau.data_ptr(), gain.data_ptr(), 1, 1, float(fs), n, 0,
                                  out.data_ptr(), wp, wl, _s._stream(dev)), "pal_render_scene")
    return out[0].double().cpu().numpy()


def _rows(signal):
    x = np.asarray(signal, dtype=np.float64)
    return x, torch.from_numpy(np.ascontiguousarray(x.reshape(1, -1).astype(np.float32))).to(_s._dev())


def normalize_signal(signal: np.ndarray) -> np.ndarray:
    """signal_processing.py:82-86."""
    x, r = _rows(signal)
    if x.size == 0:
        raise ValueError("zero-size array to reduction operation maximum which has no identity")
    return _s.normalise_compress(r, compress=False)[0].double().cpu().numpy().reshape(x.shape)


def dynamic_range_compression(signal: np.ndarray, threshold: float = 0.8, epsilon: float = 1e-8) -> np.ndarray:
    """signal_processing.py:88-94."""
    x, r = _rows(signal)
    if x.size == 0:
        raise ValueError("zero-size array to reduction operation maximum which has no identity")
    return _s.normalise_compress(r, threshold, epsilon, compress=True)[0].double().cpu().numpy().reshape(x.shape)


from .host_solver import noise_reduction  # noqa: E402,F401  (signal_processing.py:109-138, host side)
