"""The channel filter between the two stages on the GPU: signal_processing.noise_reduction's
'butterworth' branch (signal_processing.py:124-128: butter(5, [low, high], 'band') + filtfilt), applied
to every channel in main.py:191.  Host side of pal_filtfilt.

The filter DESIGN (11 + 11 coefficients, lfilter_zi) is the same scipy call the reference makes and
stays on the host; the filtering itself -- the forward / backward recursion over every sample of every
channel -- runs on the device, in float64 and in scipy's own evaluation order, so a float64 input
gives a bit-identical output.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Tuple

import numpy as np
import torch

from . import _lib


def design_butter_bandpass(fs: float, lowcut: float = 300, highcut: float = 3400, order: int = 5) -> Tuple[np.ndarray, np.ndarray]:
    """signal_processing.py:121-127: butter(5, [lowcut/nyquist, highcut/nyquist], btype='band')."""
    from scipy.signal import butter
    nyquist = 0.5 * fs
    return butter(order, [lowcut / nyquist, highcut / nyquist], btype='band')


def filtfilt_batched(x: torch.Tensor, b, a, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """scipy.signal.filtfilt(b, a, x, axis=-1) with its defaults for a CUDA tensor x [..., n] of float64 or
    float32 (arithmetic is float64 either way).  Returns a tensor of the same shape and dtype."""
    from scipy.signal import lfilter_zi
    if not (isinstance(x, torch.Tensor) and x.is_cuda):
        raise TypeError("x must be a CUDA tensor (there is no CPU path)")
    if x.dtype not in (torch.float32, torch.float64):
        raise TypeError("x must be float32 or float64")
    b = np.atleast_1d(np.asarray(b, dtype=np.float64))
    a = np.atleast_1d(np.asarray(a, dtype=np.float64))
    if a[0] != 1.0:
        b, a = b / a[0], a / a[0]                      # scipy's lfilter normalises by a[0] first
    ntaps = max(len(a), len(b))
    padlen = 3 * ntaps
    n = x.shape[-1]
    if n <= padlen:
        raise ValueError(f"The length of the input vector x must be greater than padlen, which is {padlen}.")
    bb = np.zeros(ntaps)
    bb[:len(b)] = b
    aa = np.zeros(ntaps)
    aa[:len(a)] = a
    zi = np.ascontiguousarray(lfilter_zi(b, a), dtype=np.float64)
    xc = x.contiguous()
    rows = xc.numel() // n
    y = torch.empty_like(xc) if out is None else out
    L = _lib.lib()
    need = C.c_size_t(0)
    _lib.check(L.pal_filtfilt_workspace(rows, n, padlen, C.byref(need)), "pal_filtfilt_workspace")
    ws = torch.empty(need.value, dtype=torch.uint8, device=x.device)
    wp = (ws.data_ptr() + 255) // 256 * 256
    dp = C.POINTER(C.c_double)
    with torch.cuda.device(x.device):
        rc = L.pal_filtfilt(xc.data_ptr(), rows, n, 1 if x.dtype == torch.float32 else 0, bb.ctypes.data_as(dp),
                            aa.ctypes.data_as(dp), zi.ctypes.data_as(dp), ntaps, padlen, y.data_ptr(), wp,
                            ws.numel() - (wp - ws.data_ptr()), torch.cuda.current_stream(x.device).cuda_stream)
    _lib.check(rc, "pal_filtfilt")
    for t in (xc, ws):
        t.record_stream(torch.cuda.current_stream(x.device))
    return y


def noise_reduction_batched(frames: torch.Tensor, fs: float, lowcut: float = 300, highcut: float = 3400) -> torch.Tensor:
    """noise_reduction(signal, fs, 'butterworth') for every channel of a CUDA tensor [..., n]."""
    b, a = design_butter_bandpass(fs, lowcut, highcut)
    return filtfilt_batched(frames, b, a)
