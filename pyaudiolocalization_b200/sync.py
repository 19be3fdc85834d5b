"""Channel alignment between the two stages on the GPU: utils.synchronize_signals_improved
(utils.py:407-457), called on every set of channels in main.py:188.  Host side of pal_sync_align /
pal_pad_rows.

Division of labour:
  device  channel energies and the reference channel (:415-416); the full cross-correlation of every
          channel with the reference (scipy.signal.correlate 'full', :418,426) as exact length-(2N-1)
          float64 transforms; arg-max of its magnitude (:427); the five correlation samples around it;
          the final zero padding (:448-457)
  host    per channel: the 0.3 * ref_peak test (:428), the five-point CubicSpline refinement on a
          100-point grid (:431-437, scipy's own call, float64), the plausibility limit (:443-446) and
          the rounding of the pad (:451) -- a few dozen flops on numbers that must be float64
There is no CPU fallback for the correlations: the functions raise when the CUDA library is missing.
"""
from __future__ import annotations

import ctypes as C
import logging
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib


def pads_from_alignment(ref_idx: int, peak_index: np.ndarray, absmax: np.ndarray, win: np.ndarray,
                        lens: Sequence[int], fs: float, use_interpolation: bool = True) -> np.ndarray:
    """Host remainder of utils.py:420-452 for ONE scene: from the device's arg-max indices, peak
    magnitudes and five-sample windows to the integer left pad of every channel."""
    from scipy.interpolate import CubicSpline
    m = len(peak_index)
    ref_peak = absmax[ref_idx]                                            # :419
    max_shift_samples = int(fs * 0.05)                                    # :421
    base_index = int(lens[ref_idx]) - 1                                   # :441
    shifts = []
    for idx in range(m):
        if idx == ref_idx:                                                # :423-425
            shifts.append(0)
            continue
        pk = int(peak_index[idx])
        full = int(lens[idx]) + int(lens[ref_idx]) - 1                    # len(corr)
        if absmax[idx] < 0.3 * ref_peak:                                  # :428
            logging.warning(f"Niedriger Korrelationspeak für Signal {idx} während Synchronisation. Setze Shift=0.")
            refined = pk
        elif use_interpolation and pk > 1 and pk < full - 2:              # :431
            indices = np.arange(pk - 2, pk + 3)
            cs = CubicSpline(indices, win[idx])
            fine = np.linspace(pk - 2, pk + 2, 100)
            refined = fine[np.argmax(np.abs(cs(fine)))]
        else:
            refined = pk
        shift = refined - base_index                                      # :442
        if abs(shift) > max_shift_samples:                                # :443
            logging.warning(f"Berechneter Shift ({shift} Samples) für Signal {idx} überschreitet plausiblen Bereich. Setze Shift=0.")
            shift = 0
        shifts.append(shift)
    lo = min(shifts)                                                      # :448
    return np.array([max(0, int(round(s - lo))) for s in shifts], dtype=np.int32)


def sync_align_device(frames: torch.Tensor, lens: Optional[torch.Tensor] = None):
    """pal_sync_align on frames [S, M, N] float64 (CUDA).  Returns device tensors
    (ref_idx [S] i32, peak_index [S, M] i32, absmax [S, M] f64, win [S, M, 5] f64)."""
    if not (isinstance(frames, torch.Tensor) and frames.is_cuda):
        raise TypeError("frames must be a CUDA tensor (there is no CPU path)")
    if frames.dim() != 3:
        raise ValueError("frames must have shape [S, M, N]")
    frames = frames.contiguous()
    if frames.dtype != torch.float64:
        frames = frames.double()
    s, m, n = frames.shape
    dev = frames.device
    if lens is not None:
        lens = lens.to(device=dev, dtype=torch.int32).contiguous()
        if tuple(lens.shape) != (s, m):
            raise ValueError("lens must have shape [S, M]")
    ref = torch.empty(s, dtype=torch.int32, device=dev)
    pk = torch.empty((s, m), dtype=torch.int32, device=dev)
    am = torch.empty((s, m), dtype=torch.float64, device=dev)
    win = torch.empty((s, m, 5), dtype=torch.float64, device=dev)
    L = _lib.lib()
    full, small = C.c_size_t(0), C.c_size_t(0)
    _lib.check(L.pal_sync_align_workspace(s, m, n, C.byref(full), C.byref(small)), "pal_sync_align_workspace")
    free = torch.cuda.mem_get_info(dev)[0]
    want = max(small.value, min(full.value, free // 2))
    ws = torch.empty(want + 256, dtype=torch.uint8, device=dev)
    wp = (ws.data_ptr() + 255) // 256 * 256
    with torch.cuda.device(dev):
        rc = L.pal_sync_align(frames.data_ptr(), s, m, n, lens.data_ptr() if lens is not None else None, ref.data_ptr(),
                              pk.data_ptr(), am.data_ptr(), win.data_ptr(), None, wp, ws.numel() - (wp - ws.data_ptr()),
                              torch.cuda.current_stream(dev).cuda_stream)
    _lib.check(rc, "pal_sync_align")
    for t in (frames, ws):
        t.record_stream(torch.cuda.current_stream(dev))
    return ref, pk, am, win


def pad_rows_device(frames: torch.Tensor, pads: torch.Tensor, n_out: int, lens: Optional[torch.Tensor] = None) -> torch.Tensor:
    """np.pad(sig, (pad, 0)) then right-pad to n_out for every row of frames [..., N] (float32 / float64)."""
    frames = frames.contiguous()
    if frames.dtype not in (torch.float32, torch.float64):
        raise TypeError("frames must be float32 or float64")
    n = frames.shape[-1]
    rows = frames.numel() // n
    pads = pads.to(device=frames.device, dtype=torch.int32).contiguous()
    out = torch.empty(frames.shape[:-1] + (int(n_out),), dtype=frames.dtype, device=frames.device)
    with torch.cuda.device(frames.device):
        rc = _lib.lib().pal_pad_rows(frames.data_ptr(), rows, n, lens.data_ptr() if lens is not None else None,
                                     pads.data_ptr(), out.data_ptr(), int(n_out), 1 if frames.dtype == torch.float32 else 0,
                                     torch.cuda.current_stream(frames.device).cuda_stream)
    _lib.check(rc, "pal_pad_rows")
    frames.record_stream(torch.cuda.current_stream(frames.device))
    return out


def synchronize_signals_batched(frames: torch.Tensor, fs: float, use_interpolation: bool = True,
                                lens: Optional[Sequence[Sequence[int]]] = None) -> Tuple[torch.Tensor, np.ndarray]:
    """synchronize_signals_improved for S independent scenes: frames [S, M, N] on the GPU (float64, or
    float32 which is up-cast exactly).  Returns (aligned [S, M, N_out] in the input dtype, pads [S, M]);
    N_out is the longest aligned row of the batch, shorter scenes are zero-padded on the right (within a
    scene this is exactly utils.py:454-456)."""
    s, m, n = frames.shape
    lens_np = np.full((s, m), n, np.int32) if lens is None else np.asarray(lens, np.int32).reshape(s, m)
    lens_dev = None if lens is None else torch.from_numpy(lens_np).to(frames.device)
    ref, pk, am, win = sync_align_device(frames, lens_dev)
    ref_h, pk_h, am_h, win_h = ref.cpu().numpy(), pk.cpu().numpy(), am.cpu().numpy(), win.cpu().numpy()
    pads = np.stack([pads_from_alignment(int(ref_h[i]), pk_h[i], am_h[i], win_h[i], lens_np[i], fs, use_interpolation)
                     for i in range(s)])
    n_out = int((pads + lens_np).max())
    out = pad_rows_device(frames.reshape(s * m, n), torch.from_numpy(pads.reshape(-1)), n_out,
                          None if lens_dev is None else lens_dev.reshape(-1))
    return out.reshape(s, m, n_out), pads


def synchronize_signals_improved(signals, fs, use_interpolation=True) -> List[np.ndarray]:
    """Drop-in for utils.synchronize_signals_improved (utils.py:407-457): list of 1-D arrays (lengths may
    differ) -> list of aligned float64 arrays of one common length."""
    sigs = [np.asarray(s, dtype=np.float64) for s in signals]
    lens = [len(s) for s in sigs]
    n = max(lens)
    host = np.zeros((1, len(sigs), n))
    for i, s_ in enumerate(sigs):
        host[0, i, :lens[i]] = s_
    dev = torch.device("cuda", torch.cuda.current_device())
    frames = torch.from_numpy(host).to(dev)
    out, pads = synchronize_signals_batched(frames, fs, use_interpolation, None if min(lens) == n else [lens])
    out = out[0].cpu().numpy()
    n_out = int(max(p + l for p, l in zip(pads[0], lens)))
    return [np.ascontiguousarray(out[i, :n_out]) for i in range(len(sigs))]
