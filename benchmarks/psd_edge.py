import os, sys
import numpy as np
sys.path.insert(0, os.getcwd())
import __graft_entry__ as ge
from oracle import ou as oou
pkg = ge.load_package(); ctx = pkg.default_context(0)
# the shape the fuzz found: n_t = 15642, wide frequency band (many bins)
for n_t, hi in [(15642, 0.49), (16384, 0.49), (16000, 0.3), (9000, 0.49), (8193, 0.49)]:
    fs = 1000.0; dt = 1 / fs
    t = dt * np.arange(1, n_t + 1)
    data = oou.generate_ou_with_oscillation([0.05, 10.0, 0.9], dt, n_t * dt, 5, 0.0, 1.0, seed=1, n_t=n_t)
    pp = [pkg.Normal(.1, .1), pkg.Normal(10, 5), pkg.Uniform(0, 1)]
    gm = pkg.one_timescale_and_osc_model(data, t, "abc", prior=pp, summary_method="psd", freqlims=(0.5, hi * fs))
    g = gm.gpu_model(ctx)
    d = g.distances(np.array([[0.05, 10.0, 0.9], [0.1, 12.0, 0.5]]), seed=1, generation=1)
    print(n_t, hi, g.kernel_info()[0], gm.data_sum_stats.size, d)
