#!/usr/bin/env python
"""comp_psd_lombscargle timing: 2000 x 5000 OU slices with a 10 % gap each (12 498 frequencies per slice)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge
from oracle import ou as oou
pkg = ge.load_package()
ctx = pkg.default_context(0)
n_s, n_t, fs = 2000, 5000, 100.0
d = oou.generate_ou_process(0.3, 1.0, 1 / fs, n_t / fs, n_s, seed=3, n_t=n_t)
rng = np.random.default_rng(0)
for r in range(n_s):
    a = int(rng.integers(0, n_t - 500)); d[r, a:a + 500] = np.nan
d = np.asfortranarray(d)          # trial index fastest: how a Julia (trials x time) matrix lies in memory
t = (1 / fs) * np.arange(1, n_t + 1)
pkg.comp_psd_lombscargle(t, d, None, 1 / fs, ctx=ctx)
n0 = ctx.launch_count
t0 = time.perf_counter(); p, f = pkg.comp_psd_lombscargle(t, d, None, 1 / fs, ctx=ctx); s = time.perf_counter() - t0
print(f"comp_psd_lombscargle {n_s} x {n_t} -> {p.shape}: {1e3 * s:.1f} ms end to end ({n_s / s:.0f} slices/s), {ctx.launch_count - n0} launches; "
      f"{n_s * n_t * 0.9 * p.shape[1] / s / 1e12:.2f} T (sample, frequency) pairs/s")
t0 = time.perf_counter(); r = pkg.acw(d, fs, acwtypes="knee", average_over_trials=False, oscillation_peak=False, ctx=ctx); s = time.perf_counter() - t0
print(f"acw(:knee) on the same data: {1e3 * s:.1f} ms, median tau {np.nanmedian(r.acw_results):.3f}")
