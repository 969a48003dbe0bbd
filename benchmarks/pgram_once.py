#!/usr/bin/env python
"""A few batches of one_timescale_model(summary_method = :psd) (DSP periodogram, nfft = 5000) - profiling target for
k_abc_pgram."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge
from oracle import ou as oou
pkg = ge.load_package()
ctx = pkg.default_context(0)
dt, n_t, n_trials = 1 / 500, 5000, 20
t = dt * np.arange(1, n_t + 1)
data = oou.generate_ou_process(0.3, 1.0, dt, n_t * dt, n_trials, seed=7, n_t=n_t)
m = pkg.one_timescale_model(data, t, "abc", summary_method="psd", prior=[pkg.Uniform(0.0, 5.0)])
g = m.gpu_model(ctx)
NP = int(os.environ.get("NP", 2960))
theta = np.random.default_rng(0).uniform(0.05, 5.0, size=(NP, 1))
print(g.kernel_info(), g.work())
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 2):
    ctx.timing_reset()
    d = g.distances(theta, seed=1, generation=1)
    n, ms = ctx.timing_get()
print("kernel ms", ms / max(n, 1), "samples/s", NP * n_trials * n_t / (ms / max(n, 1) * 1e-3), "nan", int(np.isnan(d).sum()))
