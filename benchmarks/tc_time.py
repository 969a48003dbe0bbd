#!/usr/bin/env python
"""C5-shape scoring throughput of k_abc_acf_tc for a given tuning (ITJL_TC_GROUPS etc. through benchmarks/_tuning.py)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import __graft_entry__ as ge
import _tuning
import bench
pkg = ge.load_package()
ctx = _tuning.from_env(pkg.default_context(0))
n_trials = int(os.environ.get("NTRIALS", 50))
data, t = bench.observed_data_host()
data = np.tile(data, (max(1, (n_trials + data.shape[0] - 1) // data.shape[0]), 1))[:n_trials]
model = pkg.one_timescale_model(data, t, "abc", prior=[pkg.Uniform(*bench.PRIOR)], n_lags=int(os.environ.get("NLAGS", 414)))
g = model.gpu_model(ctx)
NP = int(os.environ.get("NP", 94720))
theta = np.random.default_rng(0).uniform(*bench.PRIOR, size=(NP, 1))
import ctypes
plan = (ctypes.c_int32 * 16)()
pkg._lib.lib().itjl_tc_plan(model.n_t, model.n_lags, n_trials, 232448, 148, plan)
for _ in range(3):
    ctx.timing_reset()
    d = g.distances(theta, seed=1, generation=1)
    n, ms = ctx.timing_get()
print(g.kernel_info()[0], "groups(default plan)", plan[0], "kernel ms", ms / max(n, 1), "samples/s %.4g" % (NP * n_trials * model.n_t / (ms / max(n, 1) * 1e-3)), "nan", int(np.isnan(d).sum()), "d[:3]", d[:3])
