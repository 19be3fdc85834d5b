"""Batched position solve for scene sweeps: the host side of `pal_solve_positions`.

The reference fits one source position per `localize_sound_source` call with scipy's bounded least_squares
(main.py:246-274; residuals utils.py:384-405; box utils.py:364-382).  That path is unchanged (host_solver.py).  This
module solves the same least-squares problem for every scene of a batch on the device, so that a sweep ends in
positions instead of TDOA vectors."""
from __future__ import annotations

import ctypes as C
from typing import Optional

import numpy as np
import torch

from . import _lib


def solve_positions_batched(mic_positions, pairs, tdoa: torch.Tensor, c: float, weights=None, x0=None, bounds=None,
                            buffer: float = 5.0, max_iter: int = 100, xtol: float = 1e-12, ftol: float = 1e-12,
                            gtol: float = 1e-12):
    """mic_positions [M, 3] (shared) or [S, M, 3]; pairs [P, 2]; tdoa [S, P] float64 CUDA tensor (seconds, the
    reference's sign convention (d_j - d_i) = c * td).  Returns (positions [S, 3] float64, cost [S], iterations [S]) on
    the device.  `bounds` = (lo [S, 3], hi [S, 3]) or None for the reference's dynamic box; `x0` [S, 3] or None (array
    centroid)."""
    if not (isinstance(tdoa, torch.Tensor) and tdoa.is_cuda and tdoa.dtype == torch.float64 and tdoa.dim() == 2):
        raise TypeError("tdoa must be a [S, P] float64 CUDA tensor")
    dev = tdoa.device
    tdoa = tdoa.contiguous()
    s_n, p_n = tdoa.shape
    mics_np = np.ascontiguousarray(np.asarray(mic_positions, dtype=np.float64)) if not isinstance(mic_positions, torch.Tensor) else None
    mics = torch.as_tensor(mics_np).to(dev) if mics_np is not None else mic_positions.to(dev, torch.float64).contiguous()
    if mics.dim() not in (2, 3) or mics.shape[-1] != 3 or (mics.dim() == 3 and mics.shape[0] != s_n):
        raise ValueError("mic_positions must have shape [M, 3] or [S, M, 3]")
    m = int(mics.shape[-2])
    if isinstance(pairs, torch.Tensor) and pairs.is_cuda:
        # a device pair list (e.g. from gcc_phat.pairs_to_device, already range-checked): no host round trip per call
        from .gcc_phat import check_pairs_dev
        check_pairs_dev(pairs, m)
        pairs_dev = pairs
        if pairs_dev.shape[0] != p_n:
            raise ValueError("pairs must be [P, 2] with P == tdoa.shape[1]")
    else:
        pr = np.ascontiguousarray(np.asarray(pairs, dtype=np.int32).reshape(-1, 2))
        if len(pr) != p_n or (pr.size and (pr.min() < 0 or pr.max() >= m)):
            raise ValueError("pairs must be [P, 2] with 0 <= index < M and P == tdoa.shape[1]")
        pairs_dev = torch.from_numpy(pr).to(dev)

    def opt(t, shape, what):
        if t is None:
            return None
        t = torch.as_tensor(np.asarray(t, dtype=np.float64)).to(dev) if not isinstance(t, torch.Tensor) else t.to(dev, torch.float64)
        t = t.contiguous()
        if tuple(t.shape) != shape:
            raise ValueError(f"{what} must have shape {shape}")
        return t
    w = opt(weights, (p_n,), "weights")
    x0t = opt(x0, (s_n, 3), "x0")
    lo = hi = None
    if bounds is not None:
        lo, hi = opt(bounds[0], (s_n, 3), "bounds[0]"), opt(bounds[1], (s_n, 3), "bounds[1]")
    pos = torch.empty((s_n, 3), dtype=torch.float64, device=dev)
    cost = torch.empty((s_n,), dtype=torch.float64, device=dev)
    iters = torch.empty((s_n,), dtype=torch.int32, device=dev)
    L = _lib.lib()
    need = C.c_size_t(0)
    _lib.check(L.pal_solve_positions_workspace(p_n, C.byref(need)), "pal_solve_positions_workspace")
    ws = torch.empty(need.value, dtype=torch.uint8, device=dev)
    ptr = lambda t: t.data_ptr() if t is not None else None      # noqa: E731
    with torch.cuda.device(dev):
        rc = L.pal_solve_positions(mics.data_ptr(), 3 * m if mics.dim() == 3 else 0, m, pairs_dev.data_ptr(), p_n, tdoa.data_ptr(),
                                   ptr(w), ptr(x0t), ptr(lo), ptr(hi), s_n, float(c), float(buffer), int(max_iter), float(xtol),
                                   float(ftol), float(gtol), pos.data_ptr(), cost.data_ptr(), iters.data_ptr(), ws.data_ptr(),
                                   ws.numel(), torch.cuda.current_stream(dev).cuda_stream)
    _lib.check(rc, "pal_solve_positions")
    for t in (mics, pairs_dev, tdoa, ws, w, x0t, lo, hi):
        if t is not None:
            t.record_stream(torch.cuda.current_stream(dev))
    return pos, cost, iters
