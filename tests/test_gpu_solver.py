"""GPU test of the batched position solve (pal_solve_positions) against
scipy.optimize.least_squares on the reference's residuals (utils.py:384-405) and box (utils.py:364-382)."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


def _equations(x, mics, pairs, td, c, w):
    d = np.linalg.norm(x[None, :] - mics, axis=1)
    return ((d[pairs[:, 1]] - d[pairs[:, 0]]) - c * td) * w


def _bounds(mics, td, c, buffer=5.0):
    margin = buffer + max(np.percentile(c * np.abs(td), 75), 1.0)
    return mics.min(axis=0) - margin, mics.max(axis=0) + margin


@pytest.fixture(scope="module")
def pal():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import pyaudiolocalization_b200 as p
    return p


def test_batched_solve_matches_scipy_least_squares(pal):
    """64 random rooms x 8 mics, TDOAs consistent with a source inside the room plus 2 us of jitter, SNR-like weights:
    positions within 1e-6 m of scipy's bounded trf run to tight tolerances from the same start (array centroid), costs
    equal, and the reference's dynamic box reproduced."""
    from scipy.optimize import least_squares
    from pyaudiolocalization_b200 import solver, sweep
    rng = np.random.default_rng(7)
    s_n, m, c = 64, 8, 343.62
    src, mics, _ = sweep.random_shoebox_scenes(s_n, m, 123)
    pairs = pal.all_pairs(m)
    d = np.linalg.norm(src[:, None, :] - mics, axis=2)
    td = (d[:, pairs[:, 1]] - d[:, pairs[:, 0]]) / c + 2e-6 * rng.standard_normal((s_n, len(pairs)))
    w = rng.uniform(0.5, 1.5, size=len(pairs))
    pos, cost, iters = solver.solve_positions_batched(mics, pairs, torch.from_numpy(td).cuda(), c, weights=w)
    pos, cost, iters = pos.cpu().numpy(), cost.cpu().numpy(), iters.cpu().numpy()
    assert (iters > 0).all()
    worst = 0.0
    for s in range(s_n):
        lo, hi = _bounds(mics[s], td[s], c)
        x0 = np.clip(mics[s].mean(axis=0), lo, hi)
        ref = least_squares(_equations, x0, args=(mics[s], pairs, td[s], c, w), bounds=(lo, hi), method="trf",
                            ftol=1e-14, xtol=1e-14, gtol=1e-14, max_nfev=2000)
        if abs(ref.cost - cost[s]) <= 1e-9 * max(ref.cost, 1e-30) + 1e-18:      # same basin
            worst = max(worst, float(np.abs(ref.x - pos[s]).max()))
        else:
            assert cost[s] <= ref.cost * (1 + 1e-9), (s, cost[s], ref.cost)    # a different minimum must not be worse
    print(f"batched solve vs scipy trf: worst position difference {worst:.2e} m; "
          f"median error to the true source {np.median(np.linalg.norm(pos - src, axis=1)):.3e} m")
    assert worst <= 1e-6
    assert np.median(np.linalg.norm(pos - src, axis=1)) < 0.05


def test_solver_respects_explicit_bounds_and_shared_array(pal):
    """A shared microphone array, explicit starts and a box that excludes the unconstrained minimum: the solution sits
    on the box and is no worse than scipy's."""
    from scipy.optimize import least_squares
    from pyaudiolocalization_b200 import solver
    rng = np.random.default_rng(3)
    mics = rng.uniform([0, 0, 0], [4, 3, 2], size=(6, 3))
    pairs = pal.all_pairs(6)
    srcs = rng.uniform([0.5, 0.5, 0.3], [3.5, 2.5, 1.7], size=(5, 3))
    c = 343.0
    d = np.linalg.norm(srcs[:, None, :] - mics[None], axis=2)
    td = (d[:, pairs[:, 1]] - d[:, pairs[:, 0]]) / c
    lo = np.tile([0.0, 0.0, 0.0], (5, 1))
    hi = np.tile([4.0, 3.0, 1.0], (5, 1))            # sources above z = 1 are cut off
    x0 = np.tile([2.0, 1.5, 0.5], (5, 1))
    pos, cost, _ = solver.solve_positions_batched(mics, pairs, torch.from_numpy(td).cuda(), c, x0=x0, bounds=(lo, hi))
    pos, cost = pos.cpu().numpy(), cost.cpu().numpy()
    assert (pos >= lo - 1e-12).all() and (pos <= hi + 1e-12).all()
    for s in range(5):
        ref = least_squares(_equations, x0[s], args=(mics, pairs, td[s], c, 1.0), bounds=(lo[s], hi[s]), method="trf",
                            ftol=1e-14, xtol=1e-14, gtol=1e-14)
        assert cost[s] <= ref.cost + 1e-9
        if srcs[s, 2] <= 1.0:
            assert np.abs(pos[s] - srcs[s]).max() < 1e-6
