#!/usr/bin/env python
"""Secondary workloads of BASELINE.json (C1-C4), one JSON line each.  Not the driver's
bench (that is bench.py, config C5); these numbers feed DESIGN.md / profiles/."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402

MEASURED_HBM_GBS = 6453.1


def _tun(ctx=None):
    """ITJL_* environment switches of this script -> the context's itjl_tuning (benchmarks/_tuning.py)."""
    import os as _os, sys as _sys
    _sys.path.insert(0, _os.path.dirname(_os.path.abspath(__file__)))
    import _tuning
    import __graft_entry__ as _ge
    return _tuning.from_env(ctx or _ge.load_package().default_context())


def ou_data(pkg, tau, dt, n_t, n_trials, seed):
    return pkg.generate_ou_process(tau, 1.0, dt, n_t * dt, n_trials, seed=seed)


def bench_acw(pkg, ctx, kind, reps=5):
    fs = 100.0
    if kind == "randn":
        rng = np.random.default_rng(0)
        d = np.asfortranarray(rng.standard_normal((10000, 5000)))
    else:
        d = np.asfortranarray(ou_data(pkg, 0.5, 1 / fs, 5000, 10000, seed=3))
    types = ["acw0", "acw50", "acweuler", "tau"] + (["auc"] if kind != "randn" else [])
    pkg.acw(d, fs, acwtypes=types, ctx=ctx)                       # warm-up (allocations)
    ctx.timing_reset()
    t0 = time.perf_counter()
    for _ in range(reps):
        r = pkg.acw(d, fs, acwtypes=types, ctx=ctx)
    wall = (time.perf_counter() - t0) / reps
    n, ms = ctx.timing_get()
    scan_ms = ms / max(n, 1)
    gbs = d.nbytes / (scan_ms * 1e-3) / 1e9
    return {"workload": f"C2 acw() {kind} 10000x5000 fs=100 types={types}", "n_lags": r.n_lags,
            "e2e_ms": wall * 1e3, "e2e_trials_per_s": 10000 / wall, "stream_pass_ms": scan_ms,
            "stream_pass_kernels": "k_acw_stream + k_acw_finish (CUDA events on the context stream)",
            "stream_pass_GBs": gbs, "hbm_frac_of_measured": gbs / MEASURED_HBM_GBS,
            "algorithmic_bytes": int(d.nbytes)}


def bench_distances(pkg, ctx, model, n_particles, label, reps=3):
    gm = model.gpu_model(_tun(ctx))
    fl, spp = gm.work()
    rng = np.random.default_rng(1)
    theta = np.stack([np.asarray(p.rand(rng, n_particles)) for p in model.prior], axis=1)
    if theta.shape[1] == 3:
        theta[:, 0] = np.abs(theta[:, 0]) + 1e-3
    gm.distances(theta[: min(n_particles, 512)], seed=1, generation=1)
    ctx.timing_reset()
    t0 = time.perf_counter()
    for r in range(reps):
        d = gm.distances(theta, seed=1, generation=2 + r)
    wall = (time.perf_counter() - t0) / reps
    n, ms = ctx.timing_get()
    kms = ms / max(n, 1)
    return {"workload": label, "particles": n_particles, "n_stats": gm.n_stats,
            "samples_per_s_e2e": n_particles * spp / wall, "samples_per_s_kernel": n_particles * spp / (kms * 1e-3),
            "kernel_ms": kms, "flops_per_sample": fl, "tflops_algorithmic": fl * n_particles * spp / (kms * 1e-3) / 1e12,
            "nan_frac": float(np.isnan(d).mean())}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--which", default="c1,c2,c3,c4")
    args = ap.parse_args()
    pkg = ge.load_package()
    ctx = pkg.default_context(0)
    which = args.which.split(",")
    out = []
    if "c2" in which:
        out.append(bench_acw(pkg, ctx, "randn"))
        out.append(bench_acw(pkg, ctx, "ou_tau0.5"))
    if "c3" in which:
        dt, n_t = 0.002, 5000
        data = np.array(ou_data(pkg, 0.3, dt, n_t, 100, seed=7))
        rng = np.random.default_rng(7)
        for r in range(100):
            s = rng.integers(0, n_t - 500)
            data[r, s:s + 500] = np.nan
        m = pkg.one_timescale_with_missing_model(data, dt * np.arange(1, n_t + 1), "abc", prior=[pkg.Uniform(0.0, 5.0)])
        out.append(bench_distances(pkg, ctx, m, 10000, "C3 missing-data model 100x5000, 10% NaN gaps"))
    if "c4" in which:
        dt, n_t = 1 / 500, 10000
        data = pkg.generate_ou_with_oscillation([0.05, 10.0, 0.95], dt, n_t * dt, 100, 0.0, 1.0, seed=11)
        pr = [pkg.Normal(.1, .1), pkg.Normal(10, 5), pkg.Uniform(0, 1)]
        m = pkg.one_timescale_and_osc_model(np.array(data), dt * np.arange(1, n_t + 1), "abc", prior=pr)
        out.append(bench_distances(pkg, ctx, m, 2000, "C4 osc model PSD 100x10000"))
    if "c1" in which:
        dt, n_t = 0.002, 5000
        t = dt * np.arange(1, n_t + 1)
        for trials, prior, label in [(2, None, "C1 README: 2 trials, informed prior Normal(tau_hat,20)"),
                                     (30, [pkg.Uniform(0.0, 5.0)], "C1 docs variant: 30 trials, Uniform(0,5) (reference doc: 51 s)")]:
            data = np.array(ou_data(pkg, 0.3, dt, n_t, trials, seed=123))
            m = pkg.one_timescale_model(data, t, "abc", prior=prior)
            t0 = time.perf_counter()
            res, rec = pkg.pmc_abc(m, **{**pkg.get_param_dict_abc(), "verbose": False, "show_progress": False},
                                   rng=np.random.default_rng(0), seed=1, ctx=ctx, return_record=True)
            wall = time.perf_counter() - t0
            n_part = sum(r["n_total"] for r in rec)
            out.append({"workload": label, "wall_s": wall, "generations": len(rec), "particles_scored": n_part,
                        "n_lags": int(m.n_lags), "posterior_mean": float(np.average(res.final_theta[:, 0], weights=res.final_weights)),
                        "posterior_sd": float(res.final_theta[:, 0].std()), "MAP": float(res.MAP[0]),
                        "final_epsilon": res.epsilon_history[-1]})
    for o in out:
        print(json.dumps(o), flush=True)


if __name__ == "__main__":
    main()
