#!/usr/bin/env python
"""Where does the ACF error of near-unit-root trials (|tau| >> T) come from?  C1 shape (2 x 5000), tensor-core
and CUDA-core kernels, injected fp64 noise vs device Philox noise, against the fp64 C restatement."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402
from oracle import cport, models as omodels, ou as oou  # noqa: E402

pkg = ge.load_package()


def _tun(ctx=None):
    """ITJL_* environment switches of this script -> the context's itjl_tuning (benchmarks/_tuning.py)."""
    import os as _os, sys as _sys
    _sys.path.insert(0, _os.path.dirname(_os.path.abspath(__file__)))
    import _tuning
    return _tuning.from_env(ctx or pkg.default_context())

dt, n_t, n_trials = 0.002, 5000, int(os.environ.get("TRIALS", "2"))
data = oou.generate_ou_process(0.3, 1.0, dt, n_t * dt, n_trials, seed=123, n_t=n_t)
t = dt * np.arange(1, n_t + 1)
om = omodels.one_timescale_model(data, t, prior=[omodels.Uniform(0, 5)])
taus = np.array([0.3, 2.0, 5.0, 10.0, 20.0, 50.0, 100.0, 500.0, -500.0, -50.0, -20.0, -5.0, -1.0]).reshape(-1, 1)
P = taus.shape[0]
rng = np.random.default_rng(0)
u0 = rng.standard_normal((P, n_trials)); noise = rng.standard_normal((P, n_trials, n_t))
for tc in ("1", "0"):
    os.environ["ITJL_ABC_TC"] = tc
    gm = pkg.one_timescale_model(data, t, "abc", prior=[pkg.Uniform(0, 5)]).gpu_model(_tun())
    for mode in ("injected", "philox"):
        kw = dict(u0=u0, noise=noise) if mode == "injected" else dict(seed=9, generation=1)
        d_or, s_or = cport.abc_distances(om, taus, return_stats=True, **kw)
        d, s = gm.distances(taus, return_stats=True, **kw)
        err = np.abs(s - s_or)
        print(f"TC={tc} {mode:9s} " + " ".join(f"{taus[i,0]:g}:{err[i].max():.1e}@{int(err[i].argmax())}" for i in range(P)))
