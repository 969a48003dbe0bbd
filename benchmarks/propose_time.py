import os, sys, time
import numpy as np
sys.path.insert(0, os.getcwd())
import __graft_entry__ as ge
pkg = ge.load_package(); ctx = pkg.default_context(0)
rng = np.random.default_rng(0)
for n_prev in (20000, 1000000, 1000000, 437000, 1000000):
    th = rng.uniform(0.05, 5.0, size=(n_prev, 1)); w = np.full(n_prev, 1.0 / n_prev); tsq = np.array([[4.08]])
    t0 = time.perf_counter()
    out = pkg.draw_theta_pmc_device(th, w, tsq, 1000000, seed=1, generation=2, ctx=ctx)
    print(n_prev, "ms", 1e3 * (time.perf_counter() - t0), out.mean())
