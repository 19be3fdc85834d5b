"""Scene sweep: many random scenes rendered by stage 1 (main.py:66-124) and fed to stage 2
(utils.py:121-181, looped over all pairs as main.py:202-228), scenes sharded over the ranks of one box and the
per-scene integer lag vectors all-gathered once per step (BASELINE.json configs[4]).

The reference has no batch driver (one `localize_sound_source` call is one scene); this module is the batch loop a
user of the drop-in writes around `simulate_scenes_batched` + `gcc_phat_tdoa_batched`, kept in the package so that
bench.py, the tools and the tests time and check the same code.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional, Tuple

import numpy as np
import torch
import torch.distributed as dist

from . import gcc_phat as _g
from . import scene as _scene

# The materials of BASELINE cfg4 / cfg5: the stock table of materials.py prunes every
# reflection at audio-rate `freq` (headline fact 5), so sweeps that want reflections bring their own.
SWEEP_MATERIALS = {"air": {"absorption": 0.01, "freq": 1e-6}, "wood": {"absorption": 0.05, "freq": 1e-5},
                   "metal": {"absorption": 0.1, "freq": 2e-5}, "glass": {"absorption": 0.07, "freq": 1.5e-5}}
ROOM_MATERIALS = ["wood", "metal", "glass", "wood", "wood", "metal"]      # planes x=0, x=lx, y=0, y=ly, z=0, z=lz


@dataclass
class SweepConfig:
    """cfg5 of BASELINE.json: 8 mics, 0.25 s @ 16 kHz chirp (500 Hz -> 2.5 kHz), max_reflections = 3."""
    fs: int = 16000
    duration: float = 0.25
    freq: float = 500.0
    mics: int = 8
    max_reflections: int = 3
    max_expected_delay: float = 0.05
    absorption_threshold: float = 0.01
    c: float = 343.62
    signal_type: str = "chirp"

    @property
    def pairs(self) -> int:
        return self.mics * (self.mics - 1) // 2

    @property
    def samples(self) -> int:
        return int(self.duration * self.fs)


def random_shoebox_scenes(n: int, mics: int, seed: int) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """(sources [n, 3], mics [n, M, 3], plane coefficients [n, 6, 4]) of n random shoebox rooms with dimensions
    U([3, 3, 2.5], [10, 8, 4]) m, microphones and source uniform inside with a 0.3 m wall margin."""
    rng = np.random.default_rng(seed)
    dims = rng.uniform([3, 3, 2.5], [10, 8, 4], size=(n, 3))
    mic = 0.3 + rng.uniform(size=(n, mics, 3)) * (dims[:, None, :] - 0.6)
    src = 0.3 + rng.uniform(size=(n, 3)) * (dims - 0.6)
    pl = np.zeros((n, 6, 4))
    pl[:, 0, 0] = pl[:, 1, 0] = pl[:, 2, 1] = pl[:, 3, 1] = pl[:, 4, 2] = pl[:, 5, 2] = 1.0
    pl[:, 1, 3], pl[:, 3, 3], pl[:, 5, 3] = -dims[:, 0], -dims[:, 1], -dims[:, 2]
    return src, mic, pl


def planes_as_dicts(coeff: np.ndarray):
    """One scene's [6, 4] coefficients in the reference's list-of-dicts form (main.py:66-79)."""
    return [{"plane": [float(v) for v in coeff[i]], "material": ROOM_MATERIALS[i]} for i in range(coeff.shape[0])]


class SceneSweep:
    """Per-rank driver: `step(sources, mics, planes)` renders this rank's scenes chunk by chunk, runs the batched
    GCC-PHAT / TDOA pick on the rendered channels and all-gathers the lag indices of all ranks."""

    def __init__(self, cfg: SweepConfig, scenes_per_rank: int, chunk: int = 16384, device=None, group=None,
                 gather: bool = True, keep_signals: int = 0, materials=None, solve: bool = False, parts: int = 4):
        from .signal_processing import generate_signal
        self.cfg = cfg
        self.dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.n = int(scenes_per_rank)
        self.chunk = max(1, min(int(chunk), self.n))
        self.group = group
        self.world = dist.get_world_size(group) if (dist.is_available() and dist.is_initialized()) else 1
        self.do_gather = bool(gather) and self.world > 1
        self.mats = dict(SWEEP_MATERIALS if materials is None else materials)
        self.base = torch.as_tensor(generate_signal(cfg.signal_type, cfg.fs, cfg.duration, cfg.freq).astype(np.float32)).to(self.dev)
        self.cache = _scene.RenderPlanCache()
        self.k_all = torch.empty((self.n, cfg.pairs, 1), dtype=torch.int32, device=self.dev)
        self.flags = torch.empty((self.n, cfg.pairs), dtype=torch.int32, device=self.dev)
        self.gathered = torch.empty((self.world, self.n, cfg.pairs, 1), dtype=torch.int32, device=self.dev) if self.do_gather else None
        self.keep = int(keep_signals)
        self.signals: Optional[torch.Tensor] = None       # the first `keep_signals` rendered scenes of the last step
        # solve=True: the sweep ends in source positions (pal_solve_positions: the reference's residuals and box, one
        # warp per scene) instead of TDOA vectors; the lag indices are gathered either way
        self.solve = bool(solve)
        self.parts = int(parts)        # streams the grouped renderer spreads a chunk's buckets over
        self.positions = torch.empty((self.n, 3), dtype=torch.float64, device=self.dev) if self.solve else None
        self.render_ms = self.gcc_ms = self.solve_ms = 0.0
        self._ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        self.timed = False
        self.geom_stream = torch.cuda.Stream(self.dev)
        self.n_seen = (1 << 30, 0)       # smallest / largest padded length N rendered so far
        self.pairs_dev = _g.pairs_to_device(_g.all_pairs(cfg.mics), cfg.mics, self.dev)

    def warm_plans(self, n_lo: int, n_hi: int, max_streams: int = 16) -> int:
        """Build the renderer's per-length plans (chirp tables, chirp spectrum, spectrum of the base signal) for every
        padded length N in [n_lo, n_hi] ahead of time.  A plan depends on (N, base signal) only, so a long sweep pays for
        each of the few thousand possible lengths once; a short benchmark would otherwise time that one-off work.
        Returns the number of plans built."""
        n_base = self.base.numel()
        todo = [n for n in range(max(int(n_lo), n_base, 100), int(n_hi) + 1) if n not in self.cache.plans]
        if not todo:
            return 0
        self.cache._reset_for(self.base, n_base, self.dev)
        todo = [n for n in todo if n not in self.cache.plans]
        cur = torch.cuda.current_stream(self.dev)
        pool = _scene._stream_pool(self.dev, min(max_streams, len(todo)))
        ev = torch.cuda.Event()
        ev.record(cur)
        for st in pool:
            st.wait_event(ev)
        for i, n in enumerate(todo):
            self.cache.get(self.base, n_base, n, self.dev, pool[i % len(pool)])
        for st in pool:
            cur.wait_stream(st)
        return len(todo)

    def _prepare(self, sources, mics, planes, c0, c1):
        """Geometry of one chunk on the side stream (ends with the renderer's one read-back, which then waits for the
        side stream only -- not for the rendering / GCC-PHAT work of the previous chunk on the main stream)."""
        from . import main as pmain
        cfg = self.cfg
        with torch.cuda.stream(self.geom_stream):
            job = pmain.prepare_scenes_batched(sources[c0:c1], mics[c0:c1], cfg.fs, cfg.c, cfg.duration, cfg.signal_type, cfg.freq,
                                               (planes[c0:c1], ROOM_MATERIALS), self.mats, cfg.max_reflections,
                                               cfg.absorption_threshold, base_signal=self.base)
            job.done = torch.cuda.Event()
            job.done.record(self.geom_stream)
            self.n_seen = (min(self.n_seen[0], int(job.totals.min())), max(self.n_seen[1], int(job.totals.max())))
        return job

    def step(self, sources: np.ndarray, mics: np.ndarray, planes: np.ndarray):
        """One sweep step, software-pipelined over chunks: while the GPU renders chunk c and runs its GCC-PHAT stage
        (both only enqueued: neither call waits for the device), the host prepares chunk c+1 -- image sources, path
        tables, transform lengths, bucketing -- on a side stream."""
        cfg = self.cfg
        cur = torch.cuda.current_stream(self.dev)
        chunks = [(c0, min(c0 + self.chunk, self.n)) for c0 in range(0, self.n, self.chunk)]
        self.geom_stream.wait_stream(cur)
        job = self._prepare(sources, mics, planes, *chunks[0])
        for ci, (c0, c1) in enumerate(chunks):
            if self.timed:
                self._ev[0].record()
            cur.wait_event(job.done)
            sig = _scene.execute_render(job, plan_cache=self.cache, grouped_parts=self.parts)
            job_mics = job.mics
            for t in (job.tau, job.gain, job.pcount, job.src, job.mics, job.idx_dev):
                t.record_stream(cur)            # allocated under the side stream, consumed on this one
            if self.timed:
                self._ev[1].record()
            res = _g.gcc_phat_tdoa_batched(sig, float(cfg.fs), cfg.max_expected_delay, pairs_dev=self.pairs_dev)
            self.k_all[c0:c1] = res.k_idx
            self.flags[c0:c1] = res.flags
            if self.timed:
                self._ev[2].record()
            if self.solve:
                from . import solver
                td = _g.tdoa_seconds_device(res.k_idx[..., 0].contiguous(), cfg.samples, float(cfg.fs))
                # microphones and pairs are already on the device (a pageable host->device copy here would wait for
                # everything enqueued so far and stall the pipeline)
                pos, _, _ = solver.solve_positions_batched(job_mics, self.pairs_dev, td, cfg.c, max_iter=60,
                                                           xtol=1e-8, ftol=1e-8, gtol=1e-8)
                self.positions[c0:c1] = pos
            if self.keep and c0 < self.keep:
                if self.signals is None or self.signals.shape[0] != self.keep or self.signals.shape[2] != sig.shape[2]:
                    self.signals = torch.empty((self.keep,) + tuple(sig.shape[1:]), dtype=sig.dtype, device=self.dev)
                hi = min(c1, self.keep)
                self.signals[c0:hi] = sig[:hi - c0]
            if ci + 1 < len(chunks):
                job = self._prepare(sources, mics, planes, *chunks[ci + 1])       # overlaps the GPU work enqueued above
            if self.timed:
                self._ev[3].record()
                torch.cuda.synchronize()
                self.render_ms += self._ev[0].elapsed_time(self._ev[1])
                self.gcc_ms += self._ev[1].elapsed_time(self._ev[2])
                self.solve_ms += self._ev[2].elapsed_time(self._ev[3])
        if self.do_gather:
            dist.all_gather_into_tensor(self.gathered, self.k_all, group=self.group)
        return self.k_all
