#!/usr/bin/env python
"""Measured error of the oscillation models' statistics against the fp64 oracle (what the 2e-5 / PSD tolerances of the
tests have to cover): trial-mean ACF through k_abc_acf_tc and k_abc_acf, trial-mean PSD through k_abc_psd2."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge
from oracle import models as omodels, ou as oou
pkg = ge.load_package()
ctx = pkg.default_context(0)
dt, n_t, n_trials = 1 / 500, 10000, 4
t = dt * np.arange(1, n_t + 1)
data = oou.generate_ou_with_oscillation([0.05, 10.0, 0.95], dt, n_t * dt, n_trials, 0.0, 1.0, seed=11, n_t=n_t)
po = [omodels.Normal(.1, .1), omodels.Normal(10, 5), omodels.Uniform(0, 1)]
pp = [pkg.Normal(.1, .1), pkg.Normal(10, 5), pkg.Uniform(0, 1)]
theta = np.array([[0.05, 10.0, 0.95], [0.1, 12.0, 0.5], [0.02, 8.0, 0.2], [0.03, 9.5, 0.05], [0.5, 30.0, 0.99], [0.2, 3.0, 0.7]])
for sm in ("acf", "psd"):
    om = omodels.one_timescale_and_osc_model(data, t, prior=po, summary_method=sm)
    for tc in ((1, 0) if sm == "acf" else (None,)):
        if tc is not None:
            ctx.set_tuning(abc_tc=tc)
        gm = pkg.one_timescale_and_osc_model(data, t, "abc", prior=pp, summary_method=sm)
        g = gm.gpu_model(ctx)
        _, stats = g.distances(theta, seed=5, generation=3, return_stats=True)
        errs = []
        for i in range(theta.shape[0]):
            ws = omodels.summary_stats(om, omodels.generate_data(om, theta[i], seed=5, generation=3, particle=i))
            e = np.abs(stats[i] - ws)
            errs.append(e.max() if sm == "acf" else (e / ws).max())
        print(sm, g.kernel_info()[0], "max abs err (ACF) / max per-bin rel err (PSD) per particle:", " ".join(f"{e:.2e}" for e in errs))
    ctx.reset_tuning()
