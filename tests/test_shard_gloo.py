"""World-size-2/3 `gloo` tests (CPU) of the multi-GPU host logic: contiguous frame sharding and the
all-gather of per-frame lag indices.  The per-rank compute is stood in for by
the oracle (this is a test of the plumbing, not of the kernels)."""
import os
import socket

import numpy as np
import pytest

torch = pytest.importorskip("torch")
import torch.distributed as dist  # noqa: E402
import torch.multiprocessing as mp  # noqa: E402

from pyaudiolocalization_b200 import shard  # noqa: E402


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _oracle_k(frames, fs, med):
    from oracle import pal_oracle as O
    b, m, n = frames.shape
    out = np.empty((b, m * (m - 1) // 2, 1), np.int32)
    for f in range(b):
        p = 0
        for i in range(m):
            for j in range(i + 1, m):
                td, _, _ = O.get_time_delays_phat(frames[f, i], frames[f, j], fs, max_expected_delay=med)
                out[f, p, 0] = int(round(td[0] * fs)) + (n - 1)
                p += 1
    return out


def _worker(rank, world, port, total, path):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(1234)
        frames = rng.standard_normal((total, 3, 128))          # every rank draws the same batch
        lo, hi = shard.shard_range(total, world, rank)
        k_local = torch.from_numpy(_oracle_k(frames[lo:hi], 8000.0, 0.004))
        k_all = shard.gather_rows(k_local, total)
        assert k_all.shape == (total, 3, 1)
        if rank == 0:
            np.save(path, k_all.numpy())
        # every rank must hold the same gathered tensor
        chk = k_all.clone()
        dist.broadcast(chk, src=0)
        assert torch.equal(chk, k_all)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,total", [(2, 6), (2, 5), (3, 4)])
def test_sharded_gather_equals_unsharded(tmp_path, world, total):
    path = str(tmp_path / "k.npy")
    mp.spawn(_worker, args=(world, _free_port(), total, path), nprocs=world, join=True)
    got = np.load(path)
    rng = np.random.default_rng(1234)
    frames = rng.standard_normal((total, 3, 128))
    want = _oracle_k(frames, 8000.0, 0.004)
    assert np.array_equal(got, want)
    td = shard.tdoa_seconds_from_indices(got, 128, 8000.0)
    assert td.dtype == np.float64 and np.all(np.abs(td) <= 0.004 + 1e-12)


def test_shard_range_covers_everything():
    for total in (0, 1, 7, 16, 16384, 1_000_003):
        for world in (1, 2, 3, 4, 8):
            seen = 0
            for r in range(world):
                lo, hi = shard.shard_range(total, world, r)
                assert lo == min(seen, total) and hi >= lo
                seen = hi
            assert seen == total
    with pytest.raises(ValueError):
        shard.shard_range(4, 2, 2)


def test_gather_rows_without_process_group():
    x = torch.arange(12).reshape(4, 3)
    assert shard.gather_rows(x, 4) is x
    with pytest.raises(ValueError):
        shard.gather_rows(x, 5)
