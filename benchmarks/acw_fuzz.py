#!/usr/bin/env python
"""Random inputs for acw(): the GPU path against the fp64 oracle (indices bit-exact, values close) - diagnostic."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge
from oracle import acw as oacw, ou as oou
pkg = ge.load_package()
ctx = pkg.default_context(0)
rng = np.random.default_rng(int(sys.argv[1]) if len(sys.argv) > 1 else 0)
n_cases = int(sys.argv[2]) if len(sys.argv) > 2 else 40
only = int(sys.argv[3]) if len(sys.argv) > 3 else -1      # run only this case (the random stream stays the same)
bad = 0
for case in range(n_cases):
    n_slices = int(rng.choice([1, 2, 7, 31, 32, 33, 100, 129, 300]))
    n_t = int(rng.integers(40, 6000))
    fs = float(rng.choice([1.0, 100.0, 500.0]))
    gen = ["randn", "ou", "ou_offset", "smooth"][int(rng.integers(0, 4))]
    if gen == "randn":
        d = rng.standard_normal((n_slices, n_t))
    elif gen == "smooth":
        d = np.cumsum(rng.standard_normal((n_slices, n_t)), axis=1) * 0.05 + rng.standard_normal((n_slices, n_t))
    else:
        tau = float(rng.uniform(0.02, 2.0)) * (100.0 / fs)
        d = oou.generate_ou_process(tau, 1.0, 1 / fs, n_t / fs, n_slices, seed=case + 1, n_t=n_t)
        if gen == "ou_offset":
            d = d * float(rng.uniform(0.1, 10)) + float(rng.normal(0, 100))
    nan = bool(rng.integers(0, 3) == 0)
    if nan:
        for r in range(n_slices):
            if rng.integers(0, 2):
                s = int(rng.integers(0, max(1, n_t - n_t // 8))); d[r, s:s + max(1, n_t // 8)] = np.nan
    average = bool(rng.integers(0, 5) == 0)
    n_lags = None if rng.integers(0, 3) else int(rng.integers(2, max(3, n_t // 2)))
    types = ["acw0", "acw50", "acweuler", "auc", "tau"] if rng.integers(0, 2) else ["acw0", "acw50", "acweuler", "tau"]
    if only >= 0 and case != only:
        continue
    tag = f"case {case}: {gen} {n_slices}x{n_t} fs={fs} nan={nan} avg={average} n_lags={n_lags} types={len(types)}"
    want = got = None; ew = eg = None
    try:
        want = oacw.acw(d, fs, acwtypes=types, n_lags=n_lags, average_over_trials=average)
    except Exception as e:                                    # noqa: BLE001
        ew = type(e).__name__
    try:
        got = pkg.acw(d, fs, acwtypes=types, n_lags=n_lags, average_over_trials=average, ctx=ctx)
    except Exception as e:                                    # noqa: BLE001
        eg = type(e).__name__ + ": " + str(e)[:80]
    if ew or eg:
        ok = bool(ew) and bool(eg)
        bad += not ok
        print(tag, "-> oracle raised", ew, "| gpu raised", eg, "" if ok else "   <-- MISMATCH", flush=True)
        continue
    res = dict(zip(types, got.acw_results))
    msgs = []
    if got.n_lags != want["n_lags"]:
        msgs.append(f"n_lags {got.n_lags} vs {want['n_lags']}")
    for name, key in [("acw0", "k0"), ("acw50", "k50"), ("acweuler", "ke")]:
        w = np.where(want[key] >= 0, want[key] * (1.0 / fs), np.nan)   # lags = k * dt, dt = 1 / fs
        if not np.array_equal(np.atleast_1d(res[name]), w, equal_nan=True):
            msgs.append(f"{name} differs in {int(np.sum(~((np.atleast_1d(res[name]) == w) | (np.isnan(w) & np.isnan(np.atleast_1d(res[name]))))))} slices")
    if got.n_lags == want["n_lags"]:
        a = np.atleast_2d(got.acf); 
        if a.shape == want["acf"].shape:
            e = np.nanmax(np.abs(a - want["acf"])) if a.size else 0.0
            if not e < 1e-5: msgs.append(f"acf err {e:.2e}")
            if not np.array_equal(np.isnan(a), np.isnan(want["acf"])): msgs.append("acf NaN pattern differs")
        else:
            msgs.append(f"acf shape {a.shape} vs {want['acf'].shape}")
    # (an AUC can cancel to ~0: the absolute tolerance is the 1e-5 ACF tolerance integrated over the window)
    if "auc" in types and not np.allclose(np.atleast_1d(res["auc"]), want["auc"], rtol=1e-5, atol=1e-5 * max(1, int(want["k0"][0])) / fs, equal_nan=True):
        ga = np.atleast_1d(res["auc"]); rel_a = np.abs(ga - want["auc"]) / np.maximum(np.abs(want["auc"]), 1e-12)
        worst = int(np.nanargmax(rel_a))
        msgs.append(f"auc max rel err {np.nanmax(rel_a):.2e} (slice {worst}: {ga[worst]:.6g} vs {want['auc'][worst]:.6g}, window k0[0]={int(want['k0'][0])}, k0[slice]={int(want['k0'][worst])}, "
                    f"{int(np.sum(rel_a > 1e-5))} slices off)")
    if "tau" in types and not np.allclose(np.atleast_1d(res["tau"]), want["tau"], rtol=1e-2, atol=0, equal_nan=True):
        msgs.append(f"tau max rel err {np.nanmax(np.abs(np.atleast_1d(res['tau']) - want['tau']) / np.maximum(np.abs(want['tau']), 1e-12)):.2e}")
    bad += bool(msgs)
    print(tag, "-> ok" if not msgs else "-> " + "; ".join(msgs) + "   <-- MISMATCH", flush=True)
print("mismatches", bad)
