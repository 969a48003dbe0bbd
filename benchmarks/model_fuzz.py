#!/usr/bin/env python
"""Random shapes of the oscillation models (ACF and PSD summaries, with and without missing data): the fused
kernels against the fp64 C oracle on the same Philox streams (diagnostic)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge
from oracle import ou as oou, cport, models as om
pkg = ge.load_package()


def _tun(ctx=None):
    """ITJL_* environment switches of this script -> the context's itjl_tuning (benchmarks/_tuning.py)."""
    import os as _os, sys as _sys
    _sys.path.insert(0, _os.path.dirname(_os.path.abspath(__file__)))
    import _tuning
    return _tuning.from_env(ctx or pkg.default_context())

ctx = pkg.default_context(0)
rng = np.random.default_rng(int(sys.argv[1]) if len(sys.argv) > 1 else 0)
n_cases = int(sys.argv[2]) if len(sys.argv) > 2 else 30
for case in range(n_cases):
    kind = ["osc_acf", "osc_psd", "osc_acf_missing", "plain_em"][int(rng.integers(0, 4))]
    n_t = int(rng.integers(400, 12000))
    n_trials = int(rng.choice([1, 2, 3, 5, 8, 20, 50]))
    fs = float(rng.choice([250.0, 500.0, 1000.0])); dt = 1.0 / fs
    t = dt * np.arange(1, n_t + 1)
    po = [om.Normal(.1, .1), om.Normal(10, 5), om.Uniform(0, 1)]
    pp = [pkg.Normal(.1, .1), pkg.Normal(10, 5), pkg.Uniform(0, 1)]
    theta3 = np.stack([np.abs(np.random.default_rng(case).normal(.1, .05, 6)) + 0.01, np.random.default_rng(case + 1).normal(10, 3, 6),
                       np.random.default_rng(case + 2).uniform(0.05, 0.95, 6)], axis=1)
    try:
        if kind == "plain_em":
            data = oou.generate_ou_process(0.3, 1.0, dt, n_t * dt, n_trials, seed=case + 1, n_t=n_t)
            mo = om.one_timescale_model(data, t, prior=[om.Uniform(0.05, 5)], update="em")
            mg = pkg.one_timescale_model(data, t, "abc", prior=[pkg.Uniform(0.05, 5)], update="em")
            theta = np.random.default_rng(case).uniform(0.05, 5.0, size=(6, 1)); tol = 1e-5; rel = False
        else:
            dm, ds = (0.0, 1.0) if rng.integers(0, 2) else (float(rng.normal(0, 1)), float(rng.uniform(0.5, 3)))
            data = oou.generate_ou_with_oscillation([0.05, 10.0, 0.9], dt, n_t * dt, n_trials, dm, ds, seed=case + 1, n_t=n_t)
            theta = theta3
            if kind == "osc_acf_missing":
                for r in range(n_trials):
                    s = int(rng.integers(0, n_t - n_t // 10)); data[r, s:s + n_t // 10] = np.nan
                mo = om.one_timescale_and_osc_with_missing_model(data, t, prior=po)
                mg = pkg.one_timescale_and_osc_with_missing_model(data, t, "abc", prior=pp)
                tol = 2e-5; rel = False
            else:
                sm = "acf" if kind == "osc_acf" else "psd"
                mo = om.one_timescale_and_osc_model(data, t, prior=po, summary_method=sm)
                mg = pkg.one_timescale_and_osc_model(data, t, "abc", prior=pp, summary_method=sm)
                tol = 2e-5 if sm == "acf" else 1e-5; rel = sm == "psd"
        g = mg.gpu_model(_tun(ctx))
        d_or, s_or = cport.abc_distances(mo, theta, seed=4, generation=2, return_stats=True)
        d, s = g.distances(theta, seed=4, generation=2, return_stats=True)
    except Exception as e:                                    # noqa: BLE001 - diagnostic: report and go on
        print(f"case {case}: {kind} n_t={n_t} trials={n_trials} fs={fs}: {type(e).__name__}: {e}"); continue
    # PSD statistics: error relative to the spectrum's peak (a trial-mean periodogram has bins orders of magnitude
    # below it, where the fp32 FFT's rounding - relative to the peak - is all that is left); the log distance is
    # compared as well
    err = np.abs(s - s_or) / (np.max(np.abs(s_or), axis=1, keepdims=True) if rel else 1.0)
    derr = np.abs(d - d_or) / np.maximum(np.abs(d_or), 1e-30)
    flag = "" if (np.nanmax(err) <= tol and (not rel or np.nanmax(derr) <= 1e-4)) else "   <-- ABOVE TOLERANCE"
    print(f"case {case}: {kind} n_t={n_t} trials={n_trials} fs={fs} n_stats={s.shape[1]} {g.kernel_info()[0]} "
          f"max {'rel' if rel else 'abs'} stat err {np.nanmax(err):.2e} max rel dist err {np.nanmax(derr):.2e}{flag}", flush=True)
