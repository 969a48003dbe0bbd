#!/usr/bin/env python
"""A few batches of the C4 oscillation model (PSD summary) - profiling target for k_abc_psd."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge
from oracle import ou as oou
pkg = ge.load_package()


def _tun(ctx=None):
    """ITJL_* environment switches of this script -> the context's itjl_tuning (benchmarks/_tuning.py)."""
    import os as _os, sys as _sys
    _sys.path.insert(0, _os.path.dirname(_os.path.abspath(__file__)))
    import _tuning
    return _tuning.from_env(ctx or pkg.default_context())

ctx = pkg.default_context(0)
dt, n_t, n_trials = 1 / 500, 10000, 100
t = dt * np.arange(1, n_t + 1)
data = oou.generate_ou_with_oscillation([0.05, 10.0, 0.95], dt, n_t * dt, n_trials, 0.0, 1.0, seed=11, n_t=n_t)
m = pkg.one_timescale_and_osc_model(data, t, "abc", prior=[pkg.Normal(.1, .1), pkg.Normal(10, 5), pkg.Uniform(0, 1)], summary_method="psd")
g = m.gpu_model(_tun(ctx))
rng = np.random.default_rng(0)
NP = int(os.environ.get('NP', 592))
print(g.kernel_info())
theta = np.stack([np.abs(rng.normal(.1, .05, NP)) + 0.01, rng.normal(10, 3, NP), rng.uniform(0.05, 0.95, NP)], axis=1)
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 2):
    ctx.timing_reset()
    d = g.distances(theta, seed=1, generation=1)
    n, ms = ctx.timing_get()
print("kernel ms", ms / max(n, 1), "samples/s", NP * n_trials * n_t / (ms / max(n, 1) * 1e-3), "nan", int(np.isnan(d).sum()))
