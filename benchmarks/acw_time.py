#!/usr/bin/env python
"""C2 acw() with the data resident on the device: time per call and per streaming pass (CUDA events)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge
pkg = ge.load_package()
ctx = pkg.default_context(0)
d = np.asfortranarray(np.random.default_rng(7).standard_normal((10000, 5000)))
types = ["acw0", "acw50", "acweuler", "auc", "tau"]
dd = pkg.upload(d, ctx=ctx)
pkg.acw(dd, 100.0, acwtypes=types)
ctx.timing_reset()
t0 = time.perf_counter()
for _ in range(10):
    r = pkg.acw(dd, 100.0, acwtypes=types)
s = (time.perf_counter() - t0) / 10
n, ms = ctx.timing_get(); ctx.timing_stop()
print("resident ms per call", 1e3 * s, "stream + finish us", 1e3 * ms / max(n, 1), "frac of 6453 GB/s", 4e8 / (ms / max(n, 1) * 1e-3) / 6453e9, "n_lags", r.n_lags)
