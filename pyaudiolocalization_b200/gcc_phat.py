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


def pairs_to_device(pairs, m: int, dev) -> torch.Tensor:
    """Validated [P, 2] int32 device copy of a pair list (0 <= index < m), marked as checked."""
    pr = np.ascontiguousarray(np.asarray(pairs, dtype=np.int32).reshape(-1, 2))
    if pr.size and (pr.min() < 0 or pr.max() >= m):
        raise ValueError("pair index out of range")
    t = torch.from_numpy(pr).to(dev)
    t._pal_checked = (int(pr.max()) + 1 if pr.size else 0, t._version)
    return t


def check_pairs_dev(pairs_dev: torch.Tensor, m: int) -> None:
    """A caller-supplied device pair list is range-checked once per tensor object and version (one min/max
    read-back), not once per call: an out-of-range microphone index would be an out-of-bounds read in the kernels."""
    if not (pairs_dev.is_cuda and pairs_dev.dtype == torch.int32 and pairs_dev.dim() == 2 and pairs_dev.shape[1] == 2
            and pairs_dev.is_contiguous()):
        raise ValueError("pairs_dev must be a contiguous [P, 2] int32 CUDA tensor")
    seen = getattr(pairs_dev, "_pal_checked", None)
    if seen is None or seen[1] != pairs_dev._version:
        hi = -1
        if pairs_dev.numel():
            lo, hi = int(pairs_dev.min().item()), int(pairs_dev.max().item())
            if lo < 0:
                raise ValueError("pair index out of range")
        seen = (hi + 1, pairs_dev._version)
        pairs_dev._pal_checked = seen
    if seen[0] > m:
        raise ValueError("pair index out of range")


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
    n_samples: int           # n2: lag = k - (n2 - 1)
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
                          return_corr: bool = False, tie_eps: float = 1e-6, refine: bool = True,
                          workspace: Optional[torch.Tensor] = None,
                          max_workspace_bytes: Optional[int] = None,
                          out: Optional["TdoaBatch"] = None, pairs_dev: Optional[torch.Tensor] = None,
                          lengths: Optional[tuple] = None) -> TdoaBatch:
    """frames: [B, M, N] float32 CUDA tensor (float64: every row runs through the float64 kernels, the reference's own
    arithmetic type -- the path of the drop-in single-call functions).  Same options as utils.get_time_delays_phat
    (utils.py:121-127), applied to every pair of every frame.  Everything stays on the device
    and on the current CUDA stream.  N == 2048 (n = 4095) takes the fused prime-factor kernels
    and never synchronises; other lengths take the Bluestein path.  `lengths=(n1, n2)` (M == 2
    only) marks row 0 / row 1 as holding n1 / n2 valid samples (zero beyond), which is how a
    single pair of unequal signals is expressed."""
    if not (isinstance(frames, torch.Tensor) and frames.is_cuda):
        raise TypeError("frames must be a CUDA tensor (there is no CPU path)")
    if frames.dim() != 3:
        raise ValueError("frames must have shape [B, M, N]")
    frames = frames.contiguous()
    # float64 frames take the float64 ingest (pal_gcc_phat_tdoa_f64): every row through the float64 kernels, the
    # reference's arithmetic type; anything else is the float32 throughput path
    f64 = frames.dtype == torch.float64
    if not f64 and frames.dtype != torch.float32:
        frames = frames.float()
    b, m, n = frames.shape
    dev = frames.device
    if pairs_dev is None:
        pairs_dev = pairs_to_device(all_pairs(m) if pairs is None else pairs, m, dev)
    else:
        check_pairs_dev(pairs_dev, m)
    p = pairs_dev.shape[0]
    n1, n2 = (n, n) if lengths is None else (int(lengths[0]), int(lengths[1]))
    if not (1 <= n1 <= n and 1 <= n2 <= n):
        raise ValueError("lengths must satisfy 1 <= n1, n2 <= N")
    if (n1, n2) != (n, n) and m != 2:
        raise ValueError("lengths=(n1, n2) needs exactly two rows (M == 2)")
    prm = _lib.TdoaParams(window_half_width(n1, n2, fs, max_expected_delay), peak_distance(fs),
                          _METHODS.get(threshold_method, 0), float(threshold_multiplier), int(num_peaks),
                          float(tie_eps), 1 if refine else 0, n1, n2)
    if out is not None:
        k_idx, k_count, peak, gmax, flags, corr = out.k_idx, out.k_count, out.peak, out.gmax, out.flags, out.corr
        if tuple(k_idx.shape) != (b, p, num_peaks) or not all(t.is_contiguous() for t in (k_idx, k_count, peak, gmax, flags)):
            raise ValueError("`out` tensors must be contiguous with shapes [B,P,num_peaks] / [B,P]")
    else:
        k_idx = torch.empty((b, p, num_peaks), dtype=torch.int32, device=dev)
        k_count = torch.empty((b, p), dtype=torch.int32, device=dev)
        peak = torch.empty((b, p), dtype=torch.float32, device=dev)
        gmax = torch.empty((b, p), dtype=torch.float32, device=dev)
        flags = torch.empty((b, p), dtype=torch.int32, device=dev)
        corr = torch.empty((b, p, n1 + n2 - 1), dtype=torch.float32, device=dev) if return_corr else None
    if workspace is None:
        full, small = workspace_bytes(b, m, n, p)
        want = full if max_workspace_bytes is None else max(small, min(full, int(max_workspace_bytes)))
        workspace = torch.empty(want + 256, dtype=torch.uint8, device=dev)
    ws_ptr = (workspace.data_ptr() + 255) // 256 * 256
    ws_len = workspace.numel() - (ws_ptr - workspace.data_ptr())
    with torch.cuda.device(dev):
        stream = torch.cuda.current_stream(dev).cuda_stream
        entry = _lib.lib().pal_gcc_phat_tdoa_f64 if f64 else _lib.lib().pal_gcc_phat_tdoa
        rc = entry(frames.data_ptr(), b, m, n, pairs_dev.data_ptr(), p, C.byref(prm),
                                          k_idx.data_ptr(), k_count.data_ptr(), peak.data_ptr(), gmax.data_ptr(),
                                          flags.data_ptr(), corr.data_ptr() if corr is not None else None,
                                          ws_ptr, ws_len, stream)
    _lib.check(rc, "pal_gcc_phat_tdoa")
    # keep the inputs alive until the stream has consumed them
    for t in (frames, pairs_dev, workspace):
        t.record_stream(torch.cuda.current_stream(dev))
    return TdoaBatch(k_idx, k_count, peak, gmax, flags, corr, n2, float(fs))


def tdoa_seconds_device(k_idx: torch.Tensor, n_second: int, fs: float, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """time_lags[k] = (k - (n2 - 1)) / fs in float64 ON THE DEVICE (pal_tdoa_seconds): the same IEEE
    division numpy performs in utils.py:141-142, so the result is bit-identical; NaN for padding."""
    k_idx = k_idx.contiguous()
    if out is None:
        out = torch.empty(k_idx.shape, dtype=torch.float64, device=k_idx.device)
    with torch.cuda.device(k_idx.device):
        rc = _lib.lib().pal_tdoa_seconds(k_idx.data_ptr(), k_idx.numel(), int(n_second), float(fs), out.data_ptr(),
                                         torch.cuda.current_stream(k_idx.device).cuda_stream)
    _lib.check(rc, "pal_tdoa_seconds")
    return out


PCM16_SCALE = 1.0 / 32768.0     # soundfile's int16 -> float conversion (utils.py:469), exact in float32


def pcm16_to_f32(x: torch.Tensor, scale: float = PCM16_SCALE, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """int16 CUDA tensor -> float32 (x * scale) on the current stream (pal_pcm16_to_f32)."""
    if not (x.is_cuda and x.dtype == torch.int16 and x.is_contiguous()):
        raise TypeError("x must be a contiguous int16 CUDA tensor")
    if out is None:
        out = torch.empty(x.shape, dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        rc = _lib.lib().pal_pcm16_to_f32(x.data_ptr(), x.numel(), float(scale), out.data_ptr(),
                                         torch.cuda.current_stream(x.device).cuda_stream)
    _lib.check(rc, "pal_pcm16_to_f32")
    return out


def alloc_host_outputs(b: int, p: int, num_peaks: int = 1) -> dict:
    """Pinned host buffers for gcc_phat_tdoa_from_host(out_host=...).  Pinning memory is slow (tens of
    milliseconds per 100 MB), so a caller that processes batch after batch allocates them once."""
    return {"k_idx": torch.empty((b, p, num_peaks), dtype=torch.int32, pin_memory=True),
            "gmax": torch.empty((b, p), dtype=torch.float32, pin_memory=True),
            "tdoa": torch.empty((b, p, num_peaks), dtype=torch.float64, pin_memory=True)}


def gcc_phat_tdoa_from_host(frames_host: torch.Tensor, fs: float, max_expected_delay: Optional[float] = None,
                            chunk_frames: int = 256, device=None, out_host: Optional[dict] = None, **kw) -> dict:
    """End-to-end call with HOST buffers: frames_host [B, M, N] float32 on the CPU (pinned memory
    makes the copies asynchronous).  Three streams form a pipeline over chunks of frames: host->device
    copy of chunk c+1, kernels of chunk c, device->host copy of the results of chunk c-1 (lag
    indices, max(corr) and the float64 TDOA seconds, all derived on the device).  Returns numpy
    arrays backed by pinned memory.

    frames_host may also be int16 (16-bit PCM as capture hardware / WAV files deliver it): the samples cross
    PCIe as int16 -- half the bytes -- and are converted on the device to float32 x / 32768 (`pcm_scale`), which
    is exactly the float signal the reference computes on after soundfile's conversion."""
    if frames_host.is_cuda:
        raise TypeError("frames_host must live in host memory")
    pcm = frames_host.dtype == torch.int16
    pcm_scale = float(kw.pop("pcm_scale", PCM16_SCALE))
    if not pcm and frames_host.dtype != torch.float32:
        raise TypeError("frames_host must be float32 or int16")
    if kw.get("return_corr"):
        raise ValueError("return_corr is not supported by the streaming host entry point")
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    b, m, n = frames_host.shape
    num_peaks = int(kw.get("num_peaks", 1))
    pr = all_pairs(m) if kw.get("pairs") is None else np.asarray(kw.pop("pairs"), dtype=np.int32).reshape(-1, 2)
    kw.pop("pairs", None)
    p = len(pr)
    pairs_dev = pairs_to_device(pr, m, dev)
    chunk = max(1, min(int(chunk_frames), b))
    full, _ = workspace_bytes(chunk, m, n, p)
    ws = torch.empty(full + 256, dtype=torch.uint8, device=dev)
    bufs = [torch.empty((chunk, m, n), dtype=frames_host.dtype, device=dev) for _ in range(2)]
    f32buf = torch.empty((chunk, m, n), dtype=torch.float32, device=dev) if pcm else None
    res = TdoaBatch(torch.empty((b, p, num_peaks), dtype=torch.int32, device=dev),
                    torch.empty((b, p), dtype=torch.int32, device=dev),
                    torch.empty((b, p), dtype=torch.float32, device=dev),
                    torch.empty((b, p), dtype=torch.float32, device=dev),
                    torch.empty((b, p), dtype=torch.int32, device=dev), None, n, float(fs))
    td_dev = torch.empty((b, p, num_peaks), dtype=torch.float64, device=dev)
    if out_host is None:
        out_host = alloc_host_outputs(b, p, num_peaks)
    k_host, g_host, td_host = out_host["k_idx"], out_host["gmax"], out_host["tdoa"]
    if tuple(k_host.shape) != (b, p, num_peaks) or tuple(g_host.shape) != (b, p) or tuple(td_host.shape) != (b, p, num_peaks):
        raise ValueError("out_host buffers do not match [B, P, num_peaks] / [B, P]")
    copy_s, comp_s, back_s = torch.cuda.Stream(dev), torch.cuda.Stream(dev), torch.cuda.Stream(dev)
    cur = torch.cuda.current_stream(dev)
    for s_ in (copy_s, comp_s, back_s):
        s_.wait_stream(cur)
    copied = [torch.cuda.Event() for _ in range(2)]
    consumed = [torch.cuda.Event() for _ in range(2)]
    for c, f0 in enumerate(range(0, b, chunk)):
        nb = min(chunk, b - f0)
        sl = c & 1
        with torch.cuda.stream(copy_s):
            if c >= 2:
                copy_s.wait_event(consumed[sl])
            bufs[sl][:nb].copy_(frames_host[f0:f0 + nb], non_blocking=True)
            copied[sl].record(copy_s)
        with torch.cuda.stream(comp_s):
            comp_s.wait_event(copied[sl])
            view = TdoaBatch(res.k_idx[f0:f0 + nb], res.k_count[f0:f0 + nb], res.peak[f0:f0 + nb],
                             res.gmax[f0:f0 + nb], res.flags[f0:f0 + nb], None, n, float(fs))
            if pcm:
                pcm16_to_f32(bufs[sl][:nb], pcm_scale, out=f32buf[:nb])
                consumed[sl].record(comp_s)          # the int16 landing buffer is free as soon as it is converted
                gcc_phat_tdoa_batched(f32buf[:nb], fs, max_expected_delay, workspace=ws, out=view,
                                      pairs_dev=pairs_dev, **kw)
            else:
                gcc_phat_tdoa_batched(bufs[sl][:nb], fs, max_expected_delay, workspace=ws, out=view,
                                      pairs_dev=pairs_dev, **kw)
                consumed[sl].record(comp_s)
            tdoa_seconds_device(res.k_idx[f0:f0 + nb], n, fs, out=td_dev[f0:f0 + nb])
            done = torch.cuda.Event()
            done.record(comp_s)
        with torch.cuda.stream(back_s):
            back_s.wait_event(done)
            k_host[f0:f0 + nb].copy_(res.k_idx[f0:f0 + nb], non_blocking=True)
            g_host[f0:f0 + nb].copy_(res.gmax[f0:f0 + nb], non_blocking=True)
            td_host[f0:f0 + nb].copy_(td_dev[f0:f0 + nb], non_blocking=True)
    back_s.synchronize()
    cur.wait_stream(comp_s)
    cur.wait_stream(copy_s)
    return {"k_idx": k_host.numpy(), "tdoa": td_host.numpy(), "gmax": g_host.numpy(),
            "h2d_bytes": int(frames_host.numel() * frames_host.element_size()),
            "d2h_bytes": int(k_host.numel() * 4 + g_host.numel() * 4 + td_host.numel() * 8)}
