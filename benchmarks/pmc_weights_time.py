#!/usr/bin/env python
"""calc_weights at the C5 sweep's scale (one parameter, 437 048 new x 10^6 previous particles): pair sums vs the grid path."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge
pkg = ge.load_package()
ctx = pkg.default_context(0)
rng = np.random.default_rng(1)
n_prev, n = 1_000_000, 437_048
th_prev = rng.gamma(3.0, 0.8, (n_prev, 1)); w = np.full(n_prev, 1.0 / n_prev)
cov = 2.0 * np.atleast_2d(np.cov(th_prev, rowvar=False))
th = th_prev[rng.integers(0, n_prev, n)] + rng.normal(0, np.sqrt(cov[0, 0]), (n, 1))
prior = [pkg.Uniform(0.0, 50.0)]
out = {}
for mode in (1, 2, -1):
    with ctx.tuned(pmc_fp32=mode):
        pkg.calc_weights(th_prev, th, cov, w, prior, ctx=ctx)
        t0 = time.perf_counter(); out[mode] = pkg.calc_weights(th_prev, th, cov, w, prior, ctx=ctx); s = time.perf_counter() - t0
    print("pmc_fp32 =", mode, ": %.1f ms" % (1e3 * s), flush=True)
ok = out[1] > 0
print("grid vs fp32 pair sums: max relative difference", float(np.max(np.abs(out[2][ok] - out[1][ok]) / out[1][ok])))
idx = rng.integers(0, n, 200)
from oracle import abc as oabc
den = np.array([np.sum(w * oabc.mvnormal_pdf(th[i][None, :], th_prev, cov)) for i in idx])
pri = (th[idx, 0] > 0) & (th[idx, 0] < 50.0)
raw = {m: None for m in (1, 2)}
for m in (1, 2):
    with ctx.tuned(pmc_fp32=m):
        raw[m] = pkg.calc_weights(th_prev, th, cov, w, prior, ctx=ctx, normalize=False)
    rel = np.abs(raw[m][idx][pri] * den[pri] * 50.0 - 1.0)
    print("mode", m, "max relative error vs the fp64 oracle on 200 particles:", float(rel.max()))
