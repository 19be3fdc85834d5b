"""Stage 1 on the GPU: image sources (utils.py:67-106) and the multipath renderer
(main.py:66-124), host side of pal_image_sources / pal_path_table / pal_render_scene."""
from __future__ import annotations

import ctypes as C
from typing import Any, Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib


def _dev():
    if not torch.cuda.is_available():
        raise RuntimeError("pyaudiolocalization_b200 needs a CUDA device; there is no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


def _stream(dev):
    return torch.cuda.current_stream(dev).cuda_stream


_POOLS: Dict[Any, list] = {}


def _stream_pool(dev, n):
    """Side streams for launching independent buckets concurrently (created once per device)."""
    pool = _POOLS.setdefault((dev.type, dev.index), [])
    while len(pool) < n:
        pool.append(torch.cuda.Stream(dev))
    return pool[:n]


class RenderPlanCache:
    """Per-length plans of the renderer (pal_render_plan) kept on the device between calls.

    Everything the renderer derives from (N, base signal) alone -- chirp / twiddle tables, the chirp spectrum and the
    spectrum of the zero-padded base signal -- is about half of the GPU work of a small bucket, and with random rooms
    almost every scene has its own N.  A caller that renders batch after batch with the same base signal (a sweep)
    passes one cache to every call; each N is then planned once.  Plans are evicted oldest-first beyond `max_bytes`.
    Results are bit-identical with and without a cache."""

    def __init__(self, max_bytes: int = 16 << 30):
        self.max_bytes = int(max_bytes)
        self.bytes = 0
        self.plans: Dict[int, torch.Tensor] = {}
        self.key = None
        self.hits = 0
        self.misses = 0

    def _reset_for(self, base: torch.Tensor, n_base: int, dev):
        key = (base.data_ptr(), int(n_base), str(dev), int(base._version))
        if key != self.key:
            self.plans.clear()
            self.bytes = 0
            self.key = key
            self._base = base          # keeps the signal (and therefore the key's data_ptr) alive

    def get(self, base: torch.Tensor, n_base: int, total: int, dev, stream) -> Tuple[torch.Tensor, int, int]:
        """(tensor, aligned pointer, bytes) of the plan of transform length `total`, built on `stream` when missing."""
        self._reset_for(base, n_base, dev)
        t = self.plans.get(total)
        L = _lib.lib()
        need, scratch = C.c_size_t(0), C.c_size_t(0)
        _lib.check(L.pal_render_plan_bytes(int(total), C.byref(need), C.byref(scratch)), "pal_render_plan_bytes")
        if t is None:
            self.misses += 1
            while self.plans and self.bytes + need.value > self.max_bytes:
                old = next(iter(self.plans))
                self.bytes -= self.plans.pop(old).numel()
            t = torch.empty(need.value + 256, dtype=torch.uint8, device=dev)
            sc, sp, sl = _ws(scratch.value, dev)
            tp = (t.data_ptr() + 255) // 256 * 256
            _lib.check(L.pal_render_plan(base.data_ptr(), int(n_base), int(total), tp, need.value, sp, sl, stream.cuda_stream),
                       "pal_render_plan")
            sc.record_stream(stream)
            t.record_stream(stream)
            self.plans[total] = t
            self.bytes += t.numel()
        else:
            self.hits += 1
        return t, (t.data_ptr() + 255) // 256 * 256, need.value


def _ws(nbytes, dev):
    t = torch.empty(int(nbytes) + 256, dtype=torch.uint8, device=dev)
    p = (t.data_ptr() + 255) // 256 * 256
    return t, p, t.numel() - (p - t.data_ptr())


def _check_mics(mics_np: np.ndarray, n_scenes: int) -> None:
    """mic_positions is [M, 3] (shared) or [S, M, 3] (one array per scene); anything else would make the kernels read
    past the allocation (they step through it with a stride of 3*M per scene)."""
    if mics_np.ndim not in (2, 3) or mics_np.shape[-1] != 3 or mics_np.shape[-2] < 1:
        raise ValueError("mic_positions must have shape [M, 3] or [S, M, 3]")
    if mics_np.ndim == 3 and mics_np.shape[0] != n_scenes:
        raise ValueError(f"per-scene mic_positions: got {mics_np.shape[0]} arrays for {n_scenes} scenes")


class MaterialTable:
    """material_properties dict (materials.py:2-16) as index-addressed device arrays."""

    def __init__(self, mats: Dict[str, Any], dev):
        self.names = list(mats.keys())
        for n in self.names:
            if 'absorption' not in mats[n] or 'freq' not in mats[n]:
                raise ValueError(f"Absorptions- oder Frequenzeigenschaft für Material '{n}' fehlt.")
        self.index = {n: i for i, n in enumerate(self.names)}
        self.absorption = torch.tensor([float(mats[n]['absorption']) for n in self.names], dtype=torch.float64, device=dev)
        self.freq = torch.tensor([float(mats[n]['freq']) for n in self.names], dtype=torch.float64, device=dev)


def parse_planes(planes, n_scenes: int, material_index: Dict[str, int]):
    """Host-side decoding of the reflective planes of a batch (no GPU involved).

    planes: one list of {'plane': [a, b, c, d], 'material': name} dicts shared by all scenes (the reference's form,
    main.py:66-79), or a list of `n_scenes` such lists (one room per scene; same plane count and materials), or --
    the form a sweep should use, because walking tens of thousands of lists of dicts in Python costs more than
    rendering the scenes -- a pair (coefficients ndarray [n_scenes, n_planes, 4], material names [n_planes]).
    Returns (coefficients [S or 1, max(n_planes, 1), 4] float64, material ids [max(n_planes, 1)] int32, n_planes,
    per_scene).  Errors are the reference's: zero normal (utils.py:36-37), unknown material (utils.py:93-94)."""
    array_form = (isinstance(planes, (tuple, list)) and len(planes) == 2 and isinstance(planes[0], np.ndarray)
                  and planes[0].ndim == 3)
    if array_form:
        pl = np.ascontiguousarray(planes[0], dtype=np.float64)
        names = [list(planes[1])]
        if pl.shape[0] != n_scenes or pl.shape[2] != 4 or pl.shape[1] != len(names[0]):
            raise ValueError("planes array form: need coefficients [B, n_planes, 4] and n_planes material names")
        per_scene = True
    else:
        planes = list(planes or [])
        per_scene = bool(planes) and isinstance(planes[0], (list, tuple))
        plane_sets = planes if per_scene else [planes]
        if per_scene and len(plane_sets) != n_scenes:
            raise ValueError("per-scene planes: need one plane list per source")
        n0 = len(plane_sets[0])
        if any(len(pls) != n0 for pls in plane_sets):
            raise ValueError("per-scene planes: every scene needs the same number of planes")
        pl = np.array([[p['plane'] for p in pls] for pls in plane_sets], dtype=np.float64).reshape(len(plane_sets), n0, 4)
        names = [[p.get('material', 'air') for p in pls] for pls in plane_sets]
        if any(nm != names[0] for nm in names[1:]):
            raise ValueError("per-scene planes: plane materials must agree between scenes")
    n_pl = pl.shape[1]
    if n_pl and (np.einsum('spk,spk->sp', pl[:, :, :3], pl[:, :, :3]) == 0).any():
        raise ValueError("Ungültige Ebene: a^2 + b^2 + c^2 ist 0.")
    for mat in names[0]:
        if mat not in material_index:
            raise ValueError(f"Material '{mat}' ist nicht definiert. Bitte zum Dictionary hinzufügen.")
    pm = np.array([material_index[mat] for mat in names[0]], np.int32)
    if n_pl == 0:
        return np.zeros((1, 1, 4), np.float64), np.zeros(1, np.int32), 0, False
    return pl, pm, n_pl, per_scene


def image_sources_batched(sources, planes: Sequence[Dict[str, Any]], max_order: int, frequency: float,
                          material_properties: Dict[str, Any], mic_positions, absorption_threshold: float = 0.01,
                          round_decimals: int = 6, k_max: Optional[int] = None
                          ) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor, MaterialTable]:
    """Image sources of many scenes at once.  sources [B, 3]; mic_positions [M, 3] (shared) or
    [B, M, 3].  Returns device tensors (pos [B, K, 3] f64, mat [B, K] i32, count [B] i32) in the
    reference's discovery order, plus the material table that decodes `mat`."""
    dev = _dev()
    src = torch.as_tensor(np.asarray(sources, dtype=np.float64).reshape(-1, 3)).to(dev)
    b = src.shape[0]
    mics_np = np.asarray(mic_positions, dtype=np.float64)
    per_scene = mics_np.ndim == 3
    _check_mics(mics_np, b)
    mics = torch.as_tensor(np.ascontiguousarray(mics_np)).to(dev)
    n_mics = mics_np.shape[-2]
    table = MaterialTable(material_properties, dev)
    pl, pm, n_pl, per_scene_planes = parse_planes(planes, b, table.index)
    planes = [None] * n_pl
    planes_dev = torch.as_tensor(pl).to(dev)
    pm_dev = torch.as_tensor(pm).to(dev)
    if k_max is None:
        # number of distinct images is bounded by sum_o P*(P-1)^(o-1); cap the allocation
        bound, lvl = 0, len(planes)
        for _ in range(max_order):
            bound += lvl
            lvl *= max(len(planes) - 1, 1)
            if bound > 4096:
                break
        k_max = int(max(1, min(bound, 4096)))
    pos = torch.zeros((b, k_max, 3), dtype=torch.float64, device=dev)
    mat = torch.zeros((b, k_max), dtype=torch.int32, device=dev)
    cnt = torch.zeros((b,), dtype=torch.int32, device=dev)
    need = C.c_size_t(0)
    L = _lib.lib()
    _lib.check(L.pal_image_sources_workspace(len(planes), k_max, b, C.byref(need)), "pal_image_sources_workspace")
    ws, wp, wl = _ws(need.value, dev)
    rc = L.pal_image_sources(src.data_ptr(), b, planes_dev.data_ptr(), 4 * max(n_pl, 1) if per_scene_planes else 0,
                             pm_dev.data_ptr(), len(planes),
                             table.absorption.data_ptr(), table.freq.data_ptr(), mics.data_ptr(), n_mics,
                             3 * n_mics if per_scene else 0, int(max_order), float(frequency),
                             float(absorption_threshold), int(round_decimals), k_max, pos.data_ptr(), mat.data_ptr(),
                             cnt.data_ptr(), wp, wl, _stream(dev))
    _lib.check(rc, "pal_image_sources")
    return pos, mat, cnt, table


def render_scene(base_signal, source_pos, img_pos: torch.Tensor, img_mat: torch.Tensor, n_img: int, mic_positions,
                 fs: float, c: float, duration: float, freq: float, table: MaterialTable,
                 trim_to_duration: bool = True, normalise: bool = True) -> torch.Tensor:
    """main.py:94-122 for one scene; returns a [M, n_keep] float32 CUDA tensor."""
    dev = _dev()
    L = _lib.lib()
    if 'air' not in table.index:
        raise KeyError('air')                      # main.py:108 looks up material_properties['air']
    mics = torch.as_tensor(np.ascontiguousarray(np.asarray(mic_positions, dtype=np.float64).reshape(-1, 3))).to(dev)
    src = torch.as_tensor(np.asarray(source_pos, dtype=np.float64).reshape(3)).to(dev)
    m = mics.shape[0]
    k1 = int(n_img) + 1
    tau = torch.empty((m, k1), dtype=torch.float64, device=dev)
    gain = torch.empty((m, k1), dtype=torch.float64, device=dev)
    _lib.check(L.pal_path_table(src.data_ptr(), img_pos.data_ptr() if n_img else None,
                                img_mat.data_ptr() if n_img else None, int(n_img), mics.data_ptr(), m,
                                table.absorption.data_ptr(), table.freq.data_ptr(), table.index['air'], float(freq),
                                float(c), tau.data_ptr(), gain.data_ptr(), _stream(dev)), "pal_path_table")
    base = torch.as_tensor(np.ascontiguousarray(np.asarray(base_signal, dtype=np.float32))).to(dev) \
        if not isinstance(base_signal, torch.Tensor) else base_signal.to(dev, torch.float32).contiguous()
    n_base = base.numel()
    max_delay = float(tau.max().item())                         # main.py:94-101 (one small read-back)
    total = int((duration + max_delay) * fs)                    # main.py:102
    if total < n_base:
        raise ValueError("negative dimensions are not allowed")  # np.pad with a negative width
    if int(0.01 * total) < 1:
        raise ValueError("operands could not be broadcast together")   # empty fade window, signal_processing.py:78
    n_keep = min(int(duration * fs), total) if trim_to_duration else total
    out = torch.empty((m, n_keep), dtype=torch.float32, device=dev)
    full, small = C.c_size_t(0), C.c_size_t(0)
    _lib.check(L.pal_render_workspace(total, m, C.byref(full), C.byref(small)), "pal_render_workspace")
    ws, wp, wl = _ws(full.value, dev)
    _lib.check(L.pal_render_scene(base.data_ptr(), n_base, total, tau.data_ptr(), gain.data_ptr(), m, k1, float(fs),
                                  n_keep, 1 if normalise else 0, out.data_ptr(), wp, wl, _stream(dev)),
               "pal_render_scene")
    for t in (base, tau, gain, ws, mics, src):
        t.record_stream(torch.cuda.current_stream(dev))
    return out


class RenderJob:
    """Everything `execute_render` needs, produced by `prepare_render`: the per-path delay / gain tables on the device and
    the host-side bucketing of the scenes by transform length.  Preparing ends with the one read-back the renderer cannot
    avoid (N = int((duration + max delay) * fs) is formed in float64 on the host, main.py:102); executing only enqueues
    work.  A sweep prepares batch c+1 on a side stream while the GPU is busy with batch c (sweep.SceneSweep)."""
    __slots__ = ("base", "n_base", "src", "mics", "m", "tau", "gain", "pcount", "k_stride", "totals", "n_keep", "s_n", "fs",
                 "order", "uniq", "bounds", "idx_dev", "dev", "done")


def prepare_render(base_signal, sources, img_pos: torch.Tensor, img_mat: torch.Tensor, img_count: torch.Tensor, mic_positions,
                   fs: float, c: float, duration: float, freq: float, table: MaterialTable,
                   trim_to_duration: bool = True) -> RenderJob:
    """Geometry half of render_scenes_batched: path tables for all scenes in one launch, transform lengths, buckets."""
    dev = _dev()
    L = _lib.lib()
    if 'air' not in table.index:
        raise KeyError('air')
    src = torch.as_tensor(np.asarray(sources, dtype=np.float64).reshape(-1, 3)).to(dev)
    s_n = src.shape[0]
    mics_np = np.asarray(mic_positions, dtype=np.float64)
    per_scene = mics_np.ndim == 3
    _check_mics(mics_np, s_n)
    mics = torch.as_tensor(np.ascontiguousarray(mics_np)).to(dev)
    m = mics_np.shape[-2]
    for t, shape, dt, what in ((img_pos, (s_n, None, 3), torch.float64, "img_pos [S, K, 3] float64"),
                               (img_mat, (s_n, None), torch.int32, "img_mat [S, K] int32"),
                               (img_count, (s_n,), torch.int32, "img_count [S] int32")):
        if not (isinstance(t, torch.Tensor) and t.is_cuda and t.dtype == dt and t.is_contiguous() and t.dim() == len(shape)
                and all(w is None or int(g) == w for g, w in zip(t.shape, shape))):
            raise ValueError(f"render_scenes_batched: need a contiguous CUDA tensor {what} with S = {s_n} scenes")
    if img_mat.shape[1] != img_pos.shape[1]:
        raise ValueError("render_scenes_batched: img_pos and img_mat disagree on K")
    k_max = int(img_pos.shape[1])
    k_stride = k_max + 1
    tau = torch.empty((s_n, m, k_stride), dtype=torch.float64, device=dev)
    gain = torch.empty((s_n, m, k_stride), dtype=torch.float64, device=dev)
    pcount = torch.empty((s_n,), dtype=torch.int32, device=dev)
    max_tau = torch.empty((s_n,), dtype=torch.float64, device=dev)
    _lib.check(L.pal_path_table_batched(src.data_ptr(), img_pos.data_ptr(), img_mat.data_ptr(), img_count.data_ptr(), s_n,
                                        k_max, mics.data_ptr(), m, 3 * m if per_scene else 0, table.absorption.data_ptr(),
                                        table.freq.data_ptr(), table.index['air'], float(freq), float(c), k_stride,
                                        tau.data_ptr(), gain.data_ptr(), pcount.data_ptr(), max_tau.data_ptr(),
                                        _stream(dev)), "pal_path_table_batched")
    base = torch.as_tensor(np.ascontiguousarray(np.asarray(base_signal, dtype=np.float32))).to(dev) \
        if not isinstance(base_signal, torch.Tensor) else base_signal.to(dev, torch.float32).contiguous()
    n_base = base.numel()
    # ONE read-back: overflow marks of the image lists and the largest delay per scene
    stats = torch.cat([max_tau, img_count.to(torch.float64)]).cpu().numpy()
    if (stats[s_n:] < 0).any():
        raise RuntimeError("image_sources_batched overflowed k_max for some scene; call it again with a larger k_max")
    totals = ((duration + stats[:s_n]) * fs).astype(np.int64)          # main.py:102, float64 then int()
    if (totals < n_base).any():
        raise ValueError("negative dimensions are not allowed")
    if (np.floor(0.01 * totals) < 1).any():
        raise ValueError("operands could not be broadcast together")
    n_dur = int(duration * fs)
    if not trim_to_duration and len(np.unique(totals)) > 1:
        raise ValueError("trim_to_duration=False gives rows of different length; render such scenes one by one")
    job = RenderJob()
    job.base, job.n_base, job.src, job.mics, job.m = base, n_base, src, mics, m
    job.tau, job.gain, job.pcount, job.k_stride, job.totals = tau, gain, pcount, k_stride, totals
    job.n_keep = min(n_dur, int(totals.min())) if trim_to_duration else int(totals[0])
    job.s_n, job.fs, job.dev = s_n, float(fs), dev
    job.order = np.argsort(totals, kind="stable")
    job.uniq, starts = np.unique(totals[job.order], return_index=True)
    job.bounds = list(starts) + [len(job.order)]
    job.idx_dev = torch.as_tensor(job.order.astype(np.int64)).to(dev)
    return job


def execute_render(job: RenderJob, normalise: bool = True, max_workspace_bytes: int = 4 << 30, max_streams: int = 16,
                   plan_cache: Optional[RenderPlanCache] = None, grouped: bool = True, grouped_parts: int = 4) -> torch.Tensor:
    """Rendering half of render_scenes_batched: every bucket of scenes that share N is one pal_render_scenes(_planned)
    call; only enqueues work on the current stream (and a pool of side streams joined at the end)."""
    dev, L = job.dev, _lib.lib()
    base, n_base, tau, gain, pcount, k_stride, m, fs = job.base, job.n_base, job.tau, job.gain, job.pcount, job.k_stride, job.m, job.fs
    s_n, n_keep, uniq, bounds, idx_dev = job.s_n, job.n_keep, job.uniq, job.bounds, job.idx_dev
    out = torch.empty((s_n, m, n_keep), dtype=torch.float32, device=dev)
    # Scenes that share N form a bucket = one pal_render_scenes call (a few dependent launches).  With random rooms
    # almost every N is different and a bucket holds a handful of scenes, so the buckets are issued round-robin on a
    # pool of streams (the library is stateless and stream-ordered): the small grids of different buckets overlap
    # instead of queueing behind each other.  Each stream owns a workspace slice.
    if plan_cache is not None and grouped and len(uniq) > 1:
        # every bucket in ONE call (pal_render_scenes_grouped): plans first (cache misses are built here), then four
        # launches per group of buckets instead of four per bucket
        cur = torch.cuda.current_stream(dev)
        # plans the cache does not hold yet are built round-robin on the stream pool (a dozen small launches each)
        missing = [int(t) for t in uniq if int(t) not in plan_cache.plans]
        if missing:
            bpool = _stream_pool(dev, min(int(max_streams), len(missing)))
            ev0 = torch.cuda.Event()
            ev0.record(cur)
            for st in bpool:
                st.wait_event(ev0)
            for i, t in enumerate(missing):
                plan_cache.get(base, n_base, t, dev, bpool[i % len(bpool)])
            for st in bpool:
                cur.wait_stream(st)
        plans = [plan_cache.get(base, n_base, int(total), dev, cur) for total in uniq]
        n_arr = np.ascontiguousarray(uniq, dtype=np.int32)
        first = np.ascontiguousarray(bounds, dtype=np.int64)
        # The buckets go down in a few parts, each on its own stream: the transfer-function kernel of one part (FP32-bound)
        # then overlaps the convolution passes of another (shared-memory / bandwidth-bound).
        rows_cum = first * m
        n_parts = max(1, min(int(grouped_parts), len(uniq)))
        cuts = [int(np.searchsorted(rows_cum, rows_cum[-1] * k / n_parts)) for k in range(n_parts + 1)]
        cuts[0], cuts[-1] = 0, len(uniq)
        cuts = sorted(set(cuts))
        pool = _stream_pool(dev, len(cuts) - 1) if len(cuts) > 2 else [cur]
        budget = min(int(max_workspace_bytes), 3 << 30) // (len(cuts) - 1)
        ready = torch.cuda.Event()
        ready.record(cur)
        rc, held = 0, []
        for k in range(len(cuts) - 1):
            b0, b1 = cuts[k], cuts[k + 1]
            ptrs = (C.c_void_p * (b1 - b0))(*[p[1] for p in plans[b0:b1]])
            ws, wp, wl = _ws(budget, dev)
            held.append(ws)
            st = pool[k]
            if st is not cur:
                st.wait_event(ready)
            rc = L.pal_render_scenes_grouped(b1 - b0, ptrs, n_arr[b0:b1].ctypes.data, first[b0:b1 + 1].ctypes.data, tau.data_ptr(),
                                             gain.data_ptr(), pcount.data_ptr(), k_stride, idx_dev.data_ptr(), m, float(fs), n_keep,
                                             out.data_ptr(), wp, wl, st.cuda_stream)
            if rc != 0:
                break
        for st in pool:
            if st is not cur:
                cur.wait_stream(st)
        if rc == 0:
            if normalise:
                _lib.check(L.pal_normalise_compress(out.data_ptr(), s_n * m, n_keep, 0.8, 1e-8, 1, _stream(dev)),
                           "pal_normalise_compress")
            for t in [base, tau, gain, pcount, job.mics, job.src, idx_dev] + held + [p[0] for p in plans]:
                t.record_stream(cur)
            return out
        if rc != -4:                      # PAL_ERR_UNSUPPORTED: fall through to the per-bucket path
            _lib.check(rc, "pal_render_scenes_grouped")
    rows_max = int(np.max(np.diff(bounds))) * m
    need, small = C.c_size_t(0), C.c_size_t(0)
    if plan_cache is None:
        _lib.check(L.pal_render_scenes_workspace(int(uniq.max()), rows_max, C.byref(need), C.byref(small)),
                   "pal_render_scenes_workspace")
    else:
        _lib.check(L.pal_render_rows_workspace(int(uniq.max()), rows_max, C.byref(need), C.byref(small)),
                   "pal_render_rows_workspace")
    n_streams = max(1, min(int(max_streams), len(uniq), int(max_workspace_bytes) // max(int(small.value), 1)))
    per_stream = max(small.value, min(need.value, int(max_workspace_bytes) // n_streams))
    cur = torch.cuda.current_stream(dev)
    pool = [cur] if n_streams == 1 else _stream_pool(dev, n_streams)
    slices = [_ws(per_stream, dev) for _ in range(n_streams)]
    if n_streams > 1:
        ready = torch.cuda.Event()
        ready.record(cur)
        for st in pool:
            st.wait_event(ready)
    for bi, total in enumerate(uniq):
        lo, hi = int(bounds[bi]), int(bounds[bi + 1])
        k = bi % n_streams
        if plan_cache is None:
            _lib.check(L.pal_render_scenes(base.data_ptr(), n_base, int(total), tau.data_ptr(), gain.data_ptr(), pcount.data_ptr(),
                                           k_stride, idx_dev.data_ptr() + 8 * lo, hi - lo, m, float(fs), n_keep, out.data_ptr(),
                                           slices[k][1], slices[k][2], pool[k].cuda_stream), "pal_render_scenes")
        else:
            # the plan is built (first use) on the bucket's own stream; later buckets of any stream see it through
            # the end-of-call join below
            pt, pp, pb = plan_cache.get(base, n_base, int(total), dev, pool[k])
            _lib.check(L.pal_render_scenes_planned(pp, pb, n_base, int(total), tau.data_ptr(), gain.data_ptr(), pcount.data_ptr(),
                                                   k_stride, idx_dev.data_ptr() + 8 * lo, hi - lo, m, float(fs), n_keep,
                                                   out.data_ptr(), slices[k][1], slices[k][2], pool[k].cuda_stream),
                       "pal_render_scenes_planned")
            pt.record_stream(pool[k])
    if n_streams > 1:
        for st in pool:
            cur.wait_stream(st)
    ws = slices[0][0]
    for sl in slices[1:]:
        sl[0].record_stream(cur)
    if normalise:
        _lib.check(L.pal_normalise_compress(out.data_ptr(), s_n * m, n_keep, 0.8, 1e-8, 1, _stream(dev)),
                   "pal_normalise_compress")
    for t in (base, tau, gain, pcount, ws, job.mics, job.src, idx_dev):
        t.record_stream(torch.cuda.current_stream(dev))
    return out


def render_scenes_batched(base_signal, sources, img_pos: torch.Tensor, img_mat: torch.Tensor, img_count: torch.Tensor,
                          mic_positions, fs: float, c: float, duration: float, freq: float, table: MaterialTable,
                          trim_to_duration: bool = True, normalise: bool = True,
                          max_workspace_bytes: int = 4 << 30, max_streams: int = 16,
                          plan_cache: Optional[RenderPlanCache] = None) -> torch.Tensor:
    """main.py:94-122 for MANY scenes that share fs / duration / base signal: sources [S, 3],
    img_pos [S, K, 3], img_mat [S, K], img_count [S] (output of image_sources_batched),
    mic_positions [M, 3] or [S, M, 3].  Returns [S, M, n_keep] float32 on the device.

    The transform length N = int((duration + max delay) * fs) (main.py:102) differs from scene to
    scene; delays are computed for all scenes in one launch, N is formed on the host in float64
    exactly like the reference (one small read-back), and scenes that share N are rendered together.
    `plan_cache` (RenderPlanCache) keeps the per-N tables between calls that share the base signal.
    = execute_render(prepare_render(...))."""
    job = prepare_render(base_signal, sources, img_pos, img_mat, img_count, mic_positions, fs, c, duration, freq, table,
                         trim_to_duration)
    return execute_render(job, normalise, max_workspace_bytes, max_streams, plan_cache)


def normalise_compress(x: torch.Tensor, threshold: float = 0.8, epsilon: float = 1e-8, compress: bool = True):
    """In place on the rows of a float32 CUDA tensor [R, n]."""
    assert x.is_cuda and x.dtype == torch.float32 and x.is_contiguous() and x.dim() == 2
    _lib.check(_lib.lib().pal_normalise_compress(x.data_ptr(), x.shape[0], x.shape[1], float(threshold), float(epsilon),
                                                 1 if compress else 0, _stream(x.device)), "pal_normalise_compress")
    return x
