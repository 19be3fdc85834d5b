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


def test_parse_planes_host_logic():
    """scene.parse_planes (pure host code): the reference's list-of-dicts form, per-scene lists and the array form
    decode to the same coefficients / material ids; the reference's errors are kept."""
    import numpy as np
    import pytest
    from pyaudiolocalization_b200.scene import parse_planes
    idx = {"air": 0, "wood": 1, "metal": 2}
    one = [{"plane": [1, 0, 0, -5], "material": "wood"}, {"plane": [0, 1, 0, -4], "material": "metal"},
           {"plane": [0, 0, 2, -3]}]
    pl, pm, n, per = parse_planes(one, 7, idx)
    assert pl.shape == (1, 3, 4) and list(pm) == [1, 2, 0] and n == 3 and not per
    rooms = [[{"plane": [1, 0, 0, -float(s + 1)], "material": "wood"}, {"plane": [0, 1, 0, -2.0], "material": "metal"}]
             for s in range(5)]
    pl2, pm2, n2, per2 = parse_planes(rooms, 5, idx)
    assert pl2.shape == (5, 2, 4) and per2 and list(pm2) == [1, 2] and pl2[3, 0, 3] == -4.0
    pl3, pm3, n3, per3 = parse_planes((pl2.copy(), ["wood", "metal"]), 5, idx)
    assert np.array_equal(pl3, pl2) and np.array_equal(pm3, pm2) and n3 == 2 and per3
    assert parse_planes([], 3, idx)[2] == 0 and parse_planes(None, 3, idx)[2] == 0
    with pytest.raises(ValueError, match="a\\^2"):
        parse_planes([{"plane": [0, 0, 0, 1], "material": "wood"}], 1, idx)
    with pytest.raises(ValueError, match="nicht definiert"):
        parse_planes([{"plane": [1, 0, 0, 1], "material": "glass"}], 1, idx)
    with pytest.raises(ValueError):
        parse_planes(rooms, 4, idx)                                   # one room per source
    with pytest.raises(ValueError):
        parse_planes(rooms[:4] + [[rooms[4][0], {"plane": [0, 1, 0, -2.0], "material": "wood"}]], 5, idx)   # materials differ
    with pytest.raises(ValueError):
        parse_planes((pl2[:, :, :3], ["wood", "metal"]), 5, idx)
