#!/usr/bin/env python
"""Error of the fused kernels' ACF as a function of the lag for single near-unit-root trials."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge
from oracle import cport, models as omodels, ou as oou
pkg = ge.load_package()
ctx = pkg.default_context()
dt, n_t = 0.002, 5000
data = oou.generate_ou_process(0.3, 1.0, dt, n_t * dt, 1, seed=3, n_t=n_t)
t = dt * np.arange(1, n_t + 1)
om = omodels.one_timescale_model(data, t, prior=[omodels.Uniform(0, 5)], n_lags=414)
theta = np.array([[500.0]] * 8 + [[20.0]] * 8)
d_or, s_or = cport.abc_distances(om, theta, seed=31, generation=5, return_stats=True)
for tc in (1, 0):
    with ctx.tuned(abc_tc=tc):
        g = pkg.one_timescale_model(data, t, "abc", prior=[pkg.Uniform(0, 5)], n_lags=414).gpu_model(ctx)
    d, s = g.distances(theta, seed=31, generation=5, return_stats=True)
    err = s - s_or
    for i in (int(np.abs(err).max(axis=1).argmax()),):
        e = err[i]
        k = np.arange(414)
        fit = np.polyfit(1 - s_or[i], e, 1)
        print(f"tc={tc} particle {i} tau={theta[i,0]}: max|err| {np.abs(e).max():.2e}; err at lags 1,10,100,200,413: "
              + " ".join(f"{e[j]:+.2e}" for j in (1, 10, 100, 200, 413)) + f"; 1-ac at those: " + " ".join(f"{1-s_or[i][j]:.3e}" for j in (1, 10, 100, 200, 413))
              + f"; slope of err on (1-ac): {fit[0]:+.3e}, residual rms {np.std(e - np.polyval(fit, 1 - s_or[i])):.2e}")
