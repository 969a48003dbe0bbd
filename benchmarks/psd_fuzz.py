#!/usr/bin/env python
"""Random-shape comparison of the round-2 spectrum paths with the fp64 oracle: k_abc_psd2 / k_abc_psd (oscillation
model, PSD summary, C port), k_abc_pgram (periodogram model, NumPy oracle), stored-data periodogram and Lomb-Scargle,
the knee kernel.  python benchmarks/psd_fuzz.py [n_cases] [seed]"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge
from oracle import cport, knee as oknee, models as omodels, ou as oou, summary as osum
pkg = ge.load_package()
ctx = pkg.default_context(0)
n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 40
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
worst = {}
fails = 0


def note(kind, val, tol, desc):
    global fails
    worst[kind] = max(worst.get(kind, 0.0), val)
    if not val <= tol:
        fails += 1
        print("MISMATCH", kind, val, tol, desc, flush=True)


t0 = time.time()
for case in range(n_cases):
    fs = float(rng.choice([100.0, 250.0, 500.0, 1000.0]))
    dt = 1.0 / fs
    # --- oscillation model, PSD summary: both kernels over the whole range of sizes
    n_t = int(rng.integers(300, 16385))
    n_trials = int(rng.choice([1, 2, 3, 5, 8, 21]))
    t = dt * np.arange(1, n_t + 1)
    dm, ds = float(rng.normal(0, 1)), float(rng.uniform(0.5, 3))
    data = oou.generate_ou_with_oscillation([0.05, 10.0, 0.9], dt, n_t * dt, n_trials, dm, ds, seed=case + 1, n_t=n_t)
    lo = float(rng.uniform(0.1, 2.0)); hi = float(rng.uniform(0.2, 0.49)) * fs
    po = [omodels.Normal(.1, .1), omodels.Normal(10, 5), omodels.Uniform(0, 1)]
    pp = [pkg.Normal(.1, .1), pkg.Normal(10, 5), pkg.Uniform(0, 1)]
    try:
        om = omodels.one_timescale_and_osc_model(data, t, prior=po, summary_method="psd", freqlims=(lo, hi))
        gm = pkg.one_timescale_and_osc_model(data, t, "abc", prior=pp, summary_method="psd", freqlims=(lo, hi))
        theta = np.stack([np.abs(rng.normal(.1, .05, 4)) + 0.01, rng.normal(10, 3, 4), rng.uniform(0.05, 0.95, 4)], axis=1)
        d_or, s_or = cport.abc_distances(om, theta, seed=4, generation=2, return_stats=True)
        g = gm.gpu_model(ctx)
        d, s = g.distances(theta, seed=4, generation=2, return_stats=True)
        note(g.kernel_info()[0], float(np.max(np.abs(s - s_or) / np.max(np.abs(s_or), axis=1, keepdims=True))), 1e-5, (n_t, n_trials, fs))
        note(g.kernel_info()[0] + "_dist", float(np.max(np.abs(d - d_or) / d_or)), 1e-3, (n_t, n_trials, fs))
    except pkg.ItjlError as e:
        print("skipped (library refused):", n_t, n_trials, e.message, flush=True)
    # --- periodogram model
    n_t = int(rng.integers(64, 10000))
    n_trials = int(rng.choice([1, 2, 3, 4, 9]))
    t = dt * np.arange(1, n_t + 1)
    data = oou.generate_ou_process(0.05, 1.0, dt, n_t * dt, n_trials, seed=case + 7, n_t=n_t) * float(rng.uniform(0.3, 4))
    lims = (1.5 * fs / n_t, float(rng.uniform(0.2, 0.5)) * fs)
    om = omodels.one_timescale_model(data, t, summary_method="psd", prior=[omodels.Uniform(0.01, 1.0)], freqlims=lims)
    gm = pkg.one_timescale_model(data, t, "abc", summary_method="psd", prior=[pkg.Uniform(0.01, 1.0)], freqlims=lims)
    theta = rng.uniform(0.01, 1.0, size=(2, 1))
    d, s = gm.gpu_model(ctx).distances(theta, seed=6, generation=1, return_stats=True)
    for i in range(2):
        ws = omodels.summary_stats(om, omodels.generate_data(om, theta[i], seed=6, generation=1, particle=i))
        note("k_abc_pgram", float(np.max(np.abs(s[i] - ws)) / ws.max()), 1e-5, (n_t, n_trials, fs, pkg.summary.nextfastfft(n_t)))
    # --- stored data: periodogram, Lomb-Scargle
    n_sl, n_t = int(rng.integers(1, 9)), int(rng.integers(16, 9000))
    x = rng.standard_normal((n_sl, n_t)).cumsum(axis=1) * 0.05 + rng.standard_normal((n_sl, n_t)) + float(rng.normal())
    want, _ = osum.comp_psd_periodogram(x, fs)
    got, _ = pkg.summary.comp_psd(x, fs, ctx=ctx)
    note("k_pgram_data", float(np.max(np.abs(got - want) / want.max(axis=1, keepdims=True))), 1e-11, (n_sl, n_t))
    n_t = int(rng.integers(50, 1500))
    tt = dt * np.arange(1, n_t + 1)
    y = rng.standard_normal((2, n_t)).cumsum(axis=1) * 0.1 + rng.standard_normal((2, n_t))
    mask = np.zeros((2, n_t), bool)
    for r in range(2):
        a = int(rng.integers(0, n_t - n_t // 6)); mask[r, a:a + n_t // 8] = True
    y[mask] = np.nan
    want, _ = osum.comp_psd_lombscargle(tt, y, mask, dt)
    got, _ = pkg.summary.comp_psd_lombscargle(tt, y, mask, dt, ctx=ctx)
    note("k_lomb_scargle", float(np.max(np.abs(got - want))), 1e-8, (n_t,))
    # --- knee
    n = int(rng.integers(30, 3000))
    f = (np.arange(n) + 1) * float(rng.uniform(0.01, 1.0))
    amp, k = 10 ** rng.uniform(-2, 3), f[int(rng.integers(n // 20 + 1, n // 2))]
    p = amp / (1 + (f / k) ** 2) * (1 + 0.1 * rng.standard_normal(n))
    gk = pkg.utils.find_knee_frequency(p, f, ctx=ctx)
    wk = oknee.find_knee_frequency(p, f)
    obj = lambda u: float(np.mean(np.abs(oknee.lorentzian(f, u) - p)))
    # a kinked L1 objective has near-ties: two minimisers may differ by 1e-4 in the knee at 1e-8 in the objective
    note("knee_objective_excess", (obj(gk) - obj(wk)) / obj(wk), 1e-7, (n,))
    note("knee", abs(gk[1] - wk[1]) / wk[1], 1e-3, (n, gk, wk))
print(f"{n_cases} cases in {time.time() - t0:.0f} s; mismatches {fails}; worst:", {k: float(f"{v:.3g}") for k, v in worst.items()})
