#!/usr/bin/env python
"""k_knee_fit timing: find_knee_frequency (mode 0) and fooof_fit with peaks (mode 1) on batches of noisy Lorentzians."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge
pkg = ge.load_package()
ctx = pkg.default_context(0)
rng = np.random.default_rng(3)
for n_rows, n_f in ((2000, 2500), (20000, 994)):
    f = np.arange(1, n_f + 1) * 0.02
    knee = rng.uniform(0.5, 8.0, n_rows)[:, None]
    P = 3.0 / (1 + (f[None, :] / knee) ** 2) * rng.gamma(50.0, 1 / 50.0, (n_rows, n_f))      # trial-mean-like noise
    P += 0.8 * np.exp(-0.5 * ((f[None, :] - 10.0) / 0.4) ** 2)
    for name, fn in (("find_knee_frequency", lambda: pkg.find_knee_frequency(P, f, ctx=ctx)),
                     ("fooof_fit(max_peaks=1)", lambda: pkg.fooof_fit(P, f, oscillation_peak=True, max_peaks=1, return_only_knee=True, ctx=ctx))):
        fn()
        t0 = time.perf_counter(); out = fn(); s = time.perf_counter() - t0
        k = np.asarray(out)[:, 1] if name.startswith("find") else np.asarray(out)
        print(f"{name:24s} rows {n_rows:6d} n_freq {n_f:5d}: {1e3 * s:8.2f} ms  ({n_rows / s:9.0f} rows/s)  median knee err {np.median(np.abs(k - knee[:, 0]) / knee[:, 0]):.3f}", flush=True)
