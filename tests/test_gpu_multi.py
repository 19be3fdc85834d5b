"""Multi-GPU parity (run with -m gpu on a box with >= 2 GPUs; skipped otherwise): frames sharded
over 2 ranks with one NCCL all-gather of the lag indices must equal the single-GPU result bit for bit."""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
import pyaudiolocalization_b200 as pal
from pyaudiolocalization_b200 import shard
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
total = 37                                     # ragged on purpose
full = pal.synth.cfg3_frames(total, mics=6, seed=77, device="cuda")      # same seed on every rank
lo, hi = shard.shard_range(total, world, rank)
k_all, local = shard.tdoa_sharded(full[lo:hi].contiguous(), total, 16000.0, 0.05)
ref = pal.gcc_phat_tdoa_batched(full, 16000.0, 0.05)
assert torch.equal(k_all, ref.k_idx), "sharded + gathered lags differ from the single-GPU result"
td = shard.tdoa_seconds_from_indices(k_all, 2048, 16000.0)
assert np.array_equal(td, ref.tdoa_seconds())
dist.barrier()
if rank == 0:
    print("MULTI_OK", world, tuple(k_all.shape))
dist.destroy_process_group()
'''


def test_two_rank_sharding_matches_single_gpu(tmp_path):
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 CUDA devices")
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), str(script), ROOT]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert "MULTI_OK 2" in out.stdout
