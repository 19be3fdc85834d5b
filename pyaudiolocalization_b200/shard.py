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
