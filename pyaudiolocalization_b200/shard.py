"""Multi-GPU sharding of the hot path: one process per GPU, frames/scenes split in contiguous blocks,
no data-path collective; the only exchange is one all-gather of the per-frame integer lag indices
(SURVEY.md section 8e).  The reference has no counterpart (it is single-process, main.py:202-228
is one scene at a time); this is the batch driver around `gcc_phat_tdoa_batched`.

The plumbing is `torch.distributed` (NCCL on GPUs, gloo in the CPU tests).
"""
from __future__ import annotations

from typing import Optional, Tuple

import numpy as np
import torch
import torch.distributed as dist


def shard_range(total: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous block of ceil(total / world) units for `rank` (the tail ranks may get fewer or none)."""
    if total < 0 or world < 1 or not (0 <= rank < world):
        raise ValueError("shard_range: need total >= 0, world >= 1, 0 <= rank < world")
    per = -(-total // world)
    lo = min(rank * per, total)
    return lo, min(lo + per, total)


def gather_rows(local: torch.Tensor, total: int, group=None, fill=-1) -> torch.Tensor:
    """All-gather of row-sharded results.  `local` holds this rank's rows of `shard_range(total, ...)`
    (first dimension); every rank receives the full `[total, ...]` tensor.  Ragged tails are padded
    to ceil(total / world) rows for the collective and trimmed afterwards."""
    if not (dist.is_available() and dist.is_initialized()):
        if local.shape[0] != total:
            raise ValueError("gather_rows: no process group and local rows != total")
        return local
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    lo, hi = shard_range(total, world, rank)
    if local.shape[0] != hi - lo:
        raise ValueError(f"gather_rows: rank {rank} holds {local.shape[0]} rows, expected {hi - lo}")
    per = -(-total // world)
    send = local.contiguous()
    if hi - lo != per:
        pad = torch.full((per,) + tuple(local.shape[1:]), fill, dtype=local.dtype, device=local.device)
        pad[:hi - lo] = local
        send = pad
    out = torch.empty((world * per,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, send, group=group)
    return out[:total]


def tdoa_sharded(frames_local: torch.Tensor, total_frames: int, fs: float, max_expected_delay: Optional[float] = None,
                 group=None, **kw):
    """Run the batched GCC-PHAT / TDOA pick on this rank's frames and gather the integer lag indices
    of all ranks.  Returns (k_idx_all [total, P, num_peaks] int32 on the device, local TdoaBatch)."""
    from .gcc_phat import gcc_phat_tdoa_batched
    res = gcc_phat_tdoa_batched(frames_local, fs, max_expected_delay, **kw)
    return gather_rows(res.k_idx, total_frames, group=group), res


def tdoa_seconds_from_indices(k_idx, n_samples: int, fs: float) -> np.ndarray:
    """time_lags[k] = (k - (n2 - 1)) / fs in float64 from gathered integer indices (utils.py:141-142)."""
    k = (k_idx.cpu().numpy() if isinstance(k_idx, torch.Tensor) else np.asarray(k_idx)).astype(np.int64)
    td = (k - (n_samples - 1)) / float(fs)
    td[k < 0] = np.nan
    return td


class ShardedTdoa:
    """The per-rank driver of a sharded GCC-PHAT sweep: this rank's frames go through `gcc_phat_tdoa_batched`
    and the integer lag indices of all ranks are all-gathered once per step.

    Results are double-buffered: the all-gather of step i is asynchronous (NCCL's own stream) and overlaps the
    kernels of step i+1, so a rank never idles inside a step waiting for a slower peer; `drain()` waits for every
    outstanding gather.  With one rank (or no process group) there is no collective at all.
    `gathered(step)` is the [world, B_local, P, num_peaks] int32 tensor of a finished step."""

    def __init__(self, frames_per_rank: int, mics: int, n_samples: int, fs: float, max_expected_delay: Optional[float] = None,
                 pairs=None, device=None, group=None, gather: bool = True, reserve_sms: Optional[int] = None, **kw):
        from . import _lib, gcc_phat as g
        self.g = g
        self.dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.group = group
        self.world = dist.get_world_size(group) if (dist.is_available() and dist.is_initialized()) else 1
        self.rank = dist.get_rank(group) if self.world > 1 else 0
        self.fs, self.med, self.kw = float(fs), max_expected_delay, kw
        self.b, self.m, self.n = int(frames_per_rank), int(mics), int(n_samples)
        self.pairs_dev = g.pairs_to_device(g.all_pairs(mics) if pairs is None else pairs, mics, self.dev)
        self.p = int(self.pairs_dev.shape[0])
        full, _ = g.workspace_bytes(self.b, self.m, self.n, self.p)
        self.ws = torch.empty(full + 256, dtype=torch.uint8, device=self.dev)
        self.do_gather = bool(gather) and self.world > 1
        # `reserve_sms` leaves SMs out of the persistent grids for the concurrent all-gather (pal_reserve_sms).  Measured
        # on 2 x B200 (profiles/r2d_reserved_sms.md): 48.25 ms per step with 0 reserved SMs against 48.0 ms without any
        # gather and 47.95 ms on one GPU; 2 reserved SMs cost 0.7 ms, 4 cost 1.5 ms, an NCCL_MAX_CTAS cap costs up to
        # 5.6 ms.  The gather already hides behind the next step, so the default is 0.
        self.reserved = 0 if reserve_sms is None else int(reserve_sms)
        _lib.reserve_sms(self.reserved)
        nbuf = 2 if self.do_gather else 1
        self.outs = [self._new_out() for _ in range(nbuf)]
        self.bufs = [torch.empty((self.world, self.b, self.p, 1), dtype=torch.int32, device=self.dev) for _ in range(nbuf)] \
            if self.do_gather else None
        self.works = [None] * nbuf
        self.steps = 0

    def _new_out(self):
        e = lambda shape, dt: torch.empty(shape, dtype=dt, device=self.dev)      # noqa: E731
        return self.g.TdoaBatch(e((self.b, self.p, 1), torch.int32), e((self.b, self.p), torch.int32),
                                e((self.b, self.p), torch.float32), e((self.b, self.p), torch.float32),
                                e((self.b, self.p), torch.int32), None, self.n, self.fs)

    def step(self, frames_local: torch.Tensor):
        """Enqueue one step; returns this rank's TdoaBatch (valid once the stream has run)."""
        s = self.steps % len(self.outs)
        self.steps += 1
        if self.works[s] is not None:
            self.works[s].wait()
            self.works[s] = None
        out = self.g.gcc_phat_tdoa_batched(frames_local, self.fs, self.med, workspace=self.ws, out=self.outs[s],
                                           pairs_dev=self.pairs_dev, **self.kw)
        if self.do_gather:
            self.works[s] = dist.all_gather_into_tensor(self.bufs[s], out.k_idx, group=self.group, async_op=True)
        return out

    def drain(self):
        for s, w in enumerate(self.works):
            if w is not None:
                w.wait()
                self.works[s] = None

    def last(self):
        """(gathered lag indices or None, local TdoaBatch) of the most recent step, after drain()."""
        s = (self.steps - 1) % len(self.outs)
        return (self.bufs[s] if self.do_gather else None), self.outs[s]
