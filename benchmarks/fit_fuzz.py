#!/usr/bin/env python
"""Random autocorrelation rows for the batched exp-decay fit (itjl_fit_expdecay) against the oracle's argmin -
diagnostic."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge
from oracle import acwred
pkg = ge.load_package()
ctx = pkg.default_context(0)
rng = np.random.default_rng(int(sys.argv[1]) if len(sys.argv) > 1 else 0)
n_cases = int(sys.argv[2]) if len(sys.argv) > 2 else 30
bad = 0
for case in range(n_cases):
    n_lags = int(rng.integers(2, 1500)); rows = int(rng.choice([1, 3, 40]))
    dt = float(rng.choice([0.001, 0.01, 1.0]))
    lags = np.arange(n_lags) * dt
    kind = ["exp", "exp_noise", "noise", "osc", "flat", "slow"][int(rng.integers(0, 6))]
    acf = np.empty((rows, n_lags))
    for r in range(rows):
        tau = float(np.exp(rng.uniform(np.log(dt * 0.2), np.log(dt * n_lags * 5))))
        base = np.exp(-lags / tau)
        if kind == "exp": y = base
        elif kind == "exp_noise": y = base + 0.05 * rng.standard_normal(n_lags)
        elif kind == "noise": y = 0.1 * rng.standard_normal(n_lags)
        elif kind == "osc": y = base * np.cos(2 * np.pi * lags / (tau * rng.uniform(0.5, 3)))
        elif kind == "flat": y = np.full(n_lags, rng.uniform(0.2, 1.0))
        else: y = 1.0 - lags / (lags[-1] * rng.uniform(1.5, 30) + dt)
        y[0] = 1.0
        acf[r] = y
    want = np.array([acwred.fit_expdecay(lags, a) for a in acf])
    got = np.atleast_1d(pkg.fit_expdecay(lags, acf if rows > 1 else acf[0], ctx=ctx))
    rw = np.array([acwred.expdecay_residual(w, lags, a) for w, a in zip(want, acf)])
    rg = np.array([acwred.expdecay_residual(g, lags, a) if np.isfinite(g) and g > 0 else np.inf for g, a in zip(got, acf)])
    rel = np.abs(got - want) / np.abs(want)
    ok = np.all((rel <= 1e-6) | (rg <= rw * (1 + 1e-9) + 1e-15))
    bad += not ok
    print(f"case {case}: {kind} rows={rows} n_lags={n_lags} dt={dt} max rel tau err {np.nanmax(rel):.2e} worst residual ratio {np.nanmax(rg / np.maximum(rw, 1e-300)):.6f}"
          + ("" if ok else "   <-- MISMATCH"), flush=True)
print("mismatches", bad)
