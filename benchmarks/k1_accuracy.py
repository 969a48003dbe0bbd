#!/usr/bin/env python
"""Accuracy of the CUDA-core kernel k_abc_acf (and the tensor-core kernel) against the fp64 C oracle as a
function of shape (diagnostic)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge
from oracle import ou as oou, cport, models as om
pkg = ge.load_package()


def _tun(ctx=None):
    """ITJL_* environment switches of this script -> the context's itjl_tuning (benchmarks/_tuning.py)."""
    import os as _os, sys as _sys
    _sys.path.insert(0, _os.path.dirname(_os.path.abspath(__file__)))
    import _tuning
    return _tuning.from_env(ctx or pkg.default_context())

ctx = pkg.default_context(0)
dt = 0.002
for n_t, n_trials, n_lags in [(7722, 8, 100), (7722, 64, 100), (7722, 64, 400), (7722, 64, 1208), (5000, 50, 414), (5000, 200, 414), (2000, 500, 100)]:
    data = oou.generate_ou_process(0.3, 1.0, dt, n_t * dt, n_trials, seed=1, n_t=n_t)
    t = dt * np.arange(1, n_t + 1)
    m = om.one_timescale_model(data, t, prior=[om.Uniform(0.05, 5.0)], n_lags=n_lags)
    theta = np.random.default_rng(1).uniform(0.05, 5.0, size=(8, 1))
    d, s = cport.abc_distances(m, theta, seed=3, generation=1, return_stats=True)
    for tc in ("0", "1"):
        os.environ["ITJL_ABC_TC"] = tc
        g = pkg.one_timescale_model(data, t, "abc", prior=[pkg.Uniform(0.05, 5.0)], n_lags=n_lags).gpu_model(_tun(ctx))
        dg, sg = g.distances(theta, seed=3, generation=1, return_stats=True)
        e = sg - s
        print(f"n_t={n_t} trials={n_trials} L={n_lags} {g.kernel_info()[0]:13s} max|err|={np.abs(e).max():.2e} "
              f"mean err lag0={e[:, 0].mean():+.2e} lag1={e[:, 1].mean():+.2e} lag10={e[:, 10].mean():+.2e} lag50={e[:, 50].mean():+.2e}", flush=True)
