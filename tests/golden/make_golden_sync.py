#!/usr/bin/env python
"""Golden vectors for utils.synchronize_signals_improved, produced by
running the UNMODIFIED reference mounted at /root/reference in the build container.

    python tests/golden/make_golden_sync.py        ->  tests/golden/sync_vectors.npz
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from tests.golden.make_golden import import_reference  # noqa: E402


def sync_case(seed, n, m, fs, max_delay, noise, weak=None):
    rng = np.random.default_rng(seed)
    src = rng.standard_normal(n + 200)
    d = rng.integers(0, max_delay + 1, size=m)
    sigs = []
    for i in range(m):
        s = src[100 - d[i]:100 - d[i] + n] + noise * rng.standard_normal(n)
        if weak is not None and i == weak:
            s = 0.05 * rng.standard_normal(n)          # uncorrelated channel -> low-peak branch
        sigs.append(s * (1.0 + 0.1 * i))
    return sigs


def main():
    r_utils, _, _ = import_reference()
    out = {}
    cases = [(1, 1500, 4, 16000.0, 30, 0.1, None), (2, 2048, 5, 8000.0, 60, 0.3, 2), (3, 900, 3, 1000.0, 80, 0.05, None)]
    out["n_cases"] = np.array(len(cases))
    for c, (seed, n, m, fs, md, noise, weak) in enumerate(cases):
        sigs = sync_case(seed, n, m, fs, md, noise, weak)
        got = r_utils.synchronize_signals_improved([s.copy() for s in sigs], fs)
        out[f"in{c}"] = np.array(sigs)
        out[f"fs{c}"] = np.array(fs)
        out[f"out{c}"] = np.array(got)
    np.savez_compressed(os.path.join(HERE, "sync_vectors.npz"), **out)
    print("wrote", os.path.join(HERE, "sync_vectors.npz"), {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
