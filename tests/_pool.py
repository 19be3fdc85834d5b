"""Host process pool for the parity tests that push tens of thousands of rows through the oracle: workers are
SPAWNED (the pytest process already holds a CUDA context, and forking such a process is unsafe)."""
import multiprocessing as mp
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _pairs_of_frame(args):
    if ROOT not in sys.path:
        sys.path.insert(0, ROOT)
    from oracle import pal_oracle as O
    frame, fs, med = args
    sig = [np.ascontiguousarray(frame[m], dtype=np.float64) for m in range(frame.shape[0])]
    td, gm = [], []
    for i in range(len(sig)):
        for j in range(i + 1, len(sig)):
            t, corr, _ = O.get_time_delays_phat(sig[i], sig[j], fs, max_expected_delay=med)
            td.append(t[0])
            gm.append(corr.max())
    return np.array(td), np.array(gm)


def oracle_pairs(frames, fs, med, procs=None):
    """(tdoa [B, P] float64, max(corr) [B, P]) of every i<j pair of every frame, from the oracle."""
    procs = procs or min(len(frames), os.cpu_count() or 1)
    if procs <= 1:
        out = [_pairs_of_frame((f, fs, med)) for f in frames]
    else:
        with mp.get_context("spawn").Pool(procs) as pool:
            out = pool.map(_pairs_of_frame, [(f, fs, med) for f in frames], chunksize=1)
    return np.stack([o[0] for o in out]), np.stack([o[1] for o in out])
