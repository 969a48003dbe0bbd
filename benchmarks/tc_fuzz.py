#!/usr/bin/env python
"""Random shapes (n_t, n_trials, n_lags): tensor-core kernel against the CUDA-core kernel on the same Philox
streams (diagnostic; the parity tests proper are tests/test_gpu_tc.py)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge
from oracle import ou as oou

pkg = ge.load_package()


def _tun(ctx=None):
    """ITJL_* environment switches of this script -> the context's itjl_tuning (benchmarks/_tuning.py)."""
    import os as _os, sys as _sys
    _sys.path.insert(0, _os.path.dirname(_os.path.abspath(__file__)))
    import _tuning
    return _tuning.from_env(ctx or pkg.default_context())

ctx = pkg.default_context(0)
rng = np.random.default_rng(int(sys.argv[1]) if len(sys.argv) > 1 else 0)
n_cases = int(sys.argv[2]) if len(sys.argv) > 2 else 40
small = len(sys.argv) > 3 and sys.argv[3] == "small"      # short series (16 .. 299 samples)
worst = 0.0
for case in range(n_cases):
    n_t = int(rng.integers(300, 12000)) if not small else int(rng.integers(16, 300))
    n_trials = int(rng.choice([1, 2, 3, 4, 5, 7, 13, 30, 64]))
    n_lags = int(rng.integers(2, min(n_t, 1537) + 1))
    missing = bool(rng.integers(0, 4) == 0)
    dt = 0.002
    data = oou.generate_ou_process(0.3, 1.0, dt, n_t * dt, n_trials, seed=case + 1, n_t=n_t)
    if missing:
        for r in range(n_trials):
            s = int(rng.integers(0, n_t - n_t // 10)); data[r, s:s + n_t // 10] = np.nan
    t = dt * np.arange(1, n_t + 1)
    mk = (pkg.one_timescale_with_missing_model if missing else pkg.one_timescale_model)
    res = {}
    for tc in ("1", "0"):
        os.environ["ITJL_ABC_TC"] = tc
        try:
            g = mk(data, t, "abc", prior=[pkg.Uniform(0.05, 5.0)], n_lags=n_lags).gpu_model(_tun(ctx))
        except pkg.ItjlError as e:
            res[tc] = None; print(f"case {case}: n_t={n_t} trials={n_trials} lags={n_lags} missing={missing} TC={tc}: {e}"); continue
        theta = np.random.default_rng(case).uniform(0.005 if small else 0.05, 5.0, size=(8, 1))
        res[tc] = g.distances(theta, seed=3, generation=1, return_stats=True) + (g.kernel_info()[0],)
    if res["1"] is None or res["0"] is None:
        continue
    err = float(np.max(np.abs(res["1"][1] - res["0"][1])))
    worst = max(worst, err)
    flag = "" if err <= 1e-5 else "   <-- ABOVE TOLERANCE"
    if err > 1e-5:
        os.makedirs("gpurun_out", exist_ok=True)
        np.savez(f"gpurun_out/tc_fuzz_case{case}.npz", n_t=n_t, n_trials=n_trials, n_lags=n_lags, missing=missing,
                 data=data, theta=theta, s_tc=res["1"][1], s_ff=res["0"][1])
    print(f"case {case}: n_t={n_t} trials={n_trials} lags={n_lags} missing={missing} {res['1'][2]} max|dstat|={err:.2e}{flag}", flush=True)
print("worst", worst)
