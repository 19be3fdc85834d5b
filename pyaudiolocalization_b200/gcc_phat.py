"""Batched GCC-PHAT + TDOA pick on the GPU: the host side of `pal_gcc_phat_tdoa`.

Replaces the double loop of the reference, main.py:202-228, which calls
utils.get_time_delays_phat (utils.py:121-181) once per microphone pair and recomputes both
channel FFTs every time.  Here all frames x all pairs go down in one call.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Optional, Sequence

import numpy as np
import torch

from . import _lib

_METHODS = {"median": 0, "adaptive": 1}


def peak_distance(fs: float) -> int:
    """utils.py:151 `int(fs * 0.001)`; scipy.signal.find_peaks rejects distance < 1."""
    d = int(fs * 0.001)
    if d < 1:
        raise ValueError("`distance` must be greater or equal to 1")
    return d


def window_half_width(n1: int, n2: int, fs: float, max_expected_delay: Optional[float]) -> int:
    """Largest |lag| in samples for which the reference's float64 test
    `abs(lag / fs) <= max_expected_delay` (utils.py:163) holds.  -1: no window; -2: empty."""
    if max_expected_delay is None:
        return -1
    if not (0.0 <= max_expected_delay):
        return -2
    big = max(n1, n2)
    m = min(int(min(max_expected_delay * fs, float(big))) + 2, big)
    while m > 0 and not (abs(np.float64(m) / fs) <= max_expected_delay):
        m -= 1
    return m


def all_pairs(m: int) -> np.ndarray:
    """(i, j) for i < j in the order of main.py:202-203."""
    return np.array([(i, j) for i in range(m) for j in range(i + 1, m)], dtype=np.int32).reshape(-1, 2)


@dataclass
class TdoaBatch:
    k_idx: torch.Tensor      # [B, P, num_peaks] int32, raw IFFT index, -1 padded
    k_count: torch.Tensor    # [B, P] int32
    peak: torch.Tensor       # [B, P] float32, corr[k_idx[...,0]]
    gmax: torch.Tensor       # [B, P] float32, max(corr)   (main.py:223)
    flags: torch.Tensor      # [B, P] int32 (bit field, see _lib.FLAG_*)
    corr: Optional[torch.Tensor]  # [B, P, n1+n2-1] float32 (FFT order) when requested
    n_samples: int
    fs: float

    def lags(self) -> torch.Tensor:
        """Integer lag of every selected peak: k - (n2 - 1)  (scipy correlation_lags, utils.py:141)."""
        return self.k_idx - (self.n_samples - 1)

    def tdoa_seconds(self) -> np.ndarray:
        """time_lags[k] exactly as the reference computes it: int64 lag / fs in float64
        (utils.py:141-142).  Host side, from the integer indices.  Padded slots give NaN."""
        k = self.k_idx.cpu().numpy().astype(np.int64)
        td = (k - (self.n_samples - 1)) / self.fs
        td[k < 0] = np.nan
        return td


def workspace_bytes(b: int, m: int, n_samples: int, p: int) -> tuple[int, int]:
    full, small = C.c_size_t(0), C.c_size_t(0)
    _lib.check(_lib.lib().pal_gcc_phat_workspace(b, m, n_samples, p, C.byref(full), C.byref(small)),
               "pal_gcc_phat_workspace")
    return int(full.value), int(small.value)


def gcc_phat_tdoa_batched(frames: torch.Tensor, fs: float, max_expected_delay: Optional[float] = None,
                          num_peaks: int = 1, threshold_method: str = "median",
                          threshold_multiplier: float = 1.0, pairs: Optional[Sequence] = None,
                          return_corr: bool = False, tie_eps: float = 2e-6, refine: bool = True,
                          workspace: Optional[torch.Tensor] = None,
                          max_workspace_bytes: Optional[int] = None) -> TdoaBatch:
    """frames: [B, M, N] float32 CUDA tensor.  Same options as utils.get_time_delays_phat
    (utils.py:121-127), applied to every pair of every frame.  Everything stays on the device
    and on the current CUDA stream; nothing synchronises."""
    if not (isinstance(frames, torch.Tensor) and frames.is_cuda):
        raise TypeError("frames must be a CUDA tensor (there is no CPU path)")
    if frames.dim() != 3:
        raise ValueError("frames must have shape [B, M, N]")
    frames = frames.contiguous()
    if frames.dtype != torch.float32:
        frames = frames.float()
    b, m, n = frames.shape
    dev = frames.device
    pr = all_pairs(m) if pairs is None else np.ascontiguousarray(np.asarray(pairs, dtype=np.int32).reshape(-1, 2))
    if pr.size and (pr.min() < 0 or pr.max() >= m):
        raise ValueError("pair index out of range")
    p = len(pr)
    pairs_dev = torch.from_numpy(pr).to(dev)
    prm = _lib.TdoaParams(window_half_width(n, n, fs, max_expected_delay), peak_distance(fs),
                          _METHODS.get(threshold_method, 0), float(threshold_multiplier), int(num_peaks),
                          float(tie_eps), 1 if refine else 0)
    k_idx = torch.empty((b, p, num_peaks), dtype=torch.int32, device=dev)
    k_count = torch.empty((b, p), dtype=torch.int32, device=dev)
    peak = torch.empty((b, p), dtype=torch.float32, device=dev)
    gmax = torch.empty((b, p), dtype=torch.float32, device=dev)
    flags = torch.empty((b, p), dtype=torch.int32, device=dev)
    corr = torch.empty((b, p, 2 * n - 1), dtype=torch.float32, device=dev) if return_corr else None
    if workspace is None:
        full, small = workspace_bytes(b, m, n, p)
        want = full if max_workspace_bytes is None else max(small, min(full, int(max_workspace_bytes)))
        workspace = torch.empty(want + 256, dtype=torch.uint8, device=dev)
    ws_ptr = (workspace.data_ptr() + 255) // 256 * 256
    ws_len = workspace.numel() - (ws_ptr - workspace.data_ptr())
    with torch.cuda.device(dev):
        stream = torch.cuda.current_stream(dev).cuda_stream
        rc = _lib.lib().pal_gcc_phat_tdoa(frames.data_ptr(), b, m, n, pairs_dev.data_ptr(), p, C.byref(prm),
                                          k_idx.data_ptr(), k_count.data_ptr(), peak.data_ptr(), gmax.data_ptr(),
                                          flags.data_ptr(), corr.data_ptr() if corr is not None else None,
                                          ws_ptr, ws_len, stream)
    _lib.check(rc, "pal_gcc_phat_tdoa")
    # keep the inputs alive until the stream has consumed them
    for t in (frames, pairs_dev, workspace):
        t.record_stream(torch.cuda.current_stream(dev))
    return TdoaBatch(k_idx, k_count, peak, gmax, flags, corr, n, float(fs))
