#!/usr/bin/env python
"""Random shapes for the stand-alone summary / generator entry points (comp_ac_fft, comp_ac_time_missing,
comp_psd_adfriendly, generate_ou_process) against the fp64 oracle - diagnostic."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge
from oracle import summary as osum, ou as oou
pkg = ge.load_package()
ctx = pkg.default_context(0)
rng = np.random.default_rng(int(sys.argv[1]) if len(sys.argv) > 1 else 0)
n_cases = int(sys.argv[2]) if len(sys.argv) > 2 else 40
bad = 0
for case in range(n_cases):
    what = ["acf", "acf_missing", "psd", "gen", "gen_osc"][int(rng.integers(0, 5))]
    n_slices = int(rng.choice([1, 2, 5, 33, 100]))
    n_t = int(rng.integers(20, 9000))
    fs = float(rng.choice([1.0, 100.0, 500.0])); dt = 1 / fs
    tag = f"case {case}: {what} {n_slices}x{n_t} fs={fs}"
    try:
        if what in ("acf", "acf_missing", "psd"):
            x = oou.generate_ou_process(float(rng.uniform(0.02, 1.0)) * 100 / fs, 1.0, dt, n_t * dt, n_slices, seed=case + 1, n_t=n_t)
            x = x * float(rng.uniform(0.1, 5)) + float(rng.normal(0, 3))
            if what == "acf":
                n_lags = None if rng.integers(0, 2) else int(rng.integers(1, n_t + 1))
                want = osum.comp_ac_fft(x, n_lags); got = pkg.comp_ac_fft(x, n_lags, ctx=ctx)
                err = np.max(np.abs(got - want)); ok = got.shape == want.shape and err < 1e-5
            elif what == "acf_missing":
                for r in range(n_slices):
                    s = int(rng.integers(0, max(1, n_t - n_t // 7))); x[r, s:s + max(1, n_t // 7)] = np.nan
                n_lags = int(rng.integers(1, min(n_t, 1500) + 1))
                want = osum.comp_ac_time_missing(x, n_lags); got = pkg.comp_ac_time_missing(x, n_lags, ctx=ctx)
                err = np.nanmax(np.abs(got - want)); ok = got.shape == want.shape and err < 1e-5
            else:
                want, wf = osum.comp_psd_adfriendly(x, fs); got, gf = pkg.comp_psd_adfriendly(x, fs, ctx=ctx)
                err = np.max(np.abs(got - want) / want.max(axis=1, keepdims=True)); ok = got.shape == want.shape and err < 1e-5 and np.allclose(gf, wf, rtol=1e-13)
        else:
            tau = float(rng.uniform(0.02, 2.0)) * 100 / fs
            u0 = rng.standard_normal(n_slices); noise = rng.standard_normal((n_slices, n_t))
            if what == "gen":
                std = bool(rng.integers(0, 2)); D = float(rng.uniform(0.5, 3))
                want = oou.generate_ou_process(tau, D, dt, n_t * dt, n_slices, standardize=std, u0=u0, noise=noise, n_t=n_t)
                got = pkg.generate_ou_process(tau, D, dt, n_t * dt, n_slices, standardize=std, u0=u0, noise=noise, ctx=ctx)
            else:
                th = [tau, float(rng.uniform(1, 40)) * fs / 500, float(rng.uniform(-0.2, 1.2))]
                ph = rng.uniform(0, 2 * np.pi, n_slices); dm = float(rng.normal(0, 2)); ds = float(rng.uniform(0.5, 3))
                want = oou.generate_ou_with_oscillation(th, dt, n_t * dt, n_slices, dm, ds, u0=u0, noise=noise, phases=ph, n_t=n_t)
                got = pkg.generate_ou_with_oscillation(th, dt, n_t * dt, n_slices, dm, ds, u0=u0, noise=noise, phases=ph, ctx=ctx)
            scale = max(1.0, float(np.nanmax(np.abs(want))))
            err = np.nanmax(np.abs(got - want)) / scale; ok = got.shape == want.shape and err < 2e-5
    except Exception as e:                                    # noqa: BLE001
        print(tag, "->", type(e).__name__, str(e)[:120]); continue
    bad += not ok
    print(tag, f"-> err {err:.2e}" + ("" if ok else "   <-- MISMATCH"), flush=True)
print("mismatches", bad)
