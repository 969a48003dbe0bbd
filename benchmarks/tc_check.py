#!/usr/bin/env python
"""Tensor-core vs FFMA fused ACF kernel on a C5-shaped model: agreement with the fp64 oracle and timing."""
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _tun(ctx=None):
    """ITJL_* environment switches of this script -> the context's itjl_tuning (benchmarks/_tuning.py)."""
    import os as _os, sys as _sys
    _sys.path.insert(0, _os.path.dirname(_os.path.abspath(__file__)))
    import _tuning
    import __graft_entry__ as _ge
    return _tuning.from_env(ctx or _ge.load_package().default_context())


def child(mode, n_particles, tag):
    os.environ["ITJL_ABC_TC"] = mode
    import __graft_entry__ as ge
    import bench
    pkg = ge.load_package()
    ctx = pkg.default_context(0)
    data, t = bench.observed_data_host()
    model = pkg.one_timescale_model(data, t, "abc", prior=[pkg.Uniform(0.05, 5.0)])
    gm = model.gpu_model(_tun(ctx))
    rng = np.random.default_rng(3)
    theta = rng.uniform(0.05, 5.0, size=(n_particles, 1))
    # oracle comparison with injected noise on 6 particles (incl. long and short timescales)
    from oracle import ou as oou, summary as osum
    th6 = np.array([[0.3], [4.9], [0.05], [1.0], [2.5], [0.1]])
    n_trials, n_t = data.shape
    u0 = rng.standard_normal((6, n_trials)); noise = rng.standard_normal((6, n_trials, n_t))
    d6, st6 = gm.distances(th6, u0=u0, noise=noise, return_stats=True)
    errs = []
    for i in range(6):
        x = oou.generate_ou_process(th6[i, 0], model.data_sd, model.dt, n_t * model.dt, n_trials, u0=u0[i], noise=noise[i], n_t=n_t)
        ac = osum.comp_ac_fft(x, model.n_lags).mean(axis=0)
        errs.append((np.max(np.abs(st6[i] - ac)), int(np.argmax(np.abs(st6[i] - ac))), float(np.mean(st6[i] - ac))))
    print(f"[{tag}] vs fp64 oracle (max abs err, argmax lag, mean signed):", [(f"{e:.2e}", k, f"{m:+.1e}") for e, k, m in errs], flush=True)
    gm.distances(theta[:296], seed=9, generation=1)
    ctx.timing_reset()
    t0 = time.perf_counter()
    d = gm.distances(theta, seed=9, generation=1)
    wall = time.perf_counter() - t0
    n, ms = ctx.timing_get()
    d2 = gm.distances(theta[:5000], seed=9, generation=1)
    print(f"[{tag}] n_lags={model.n_lags} particles={n_particles} kernel_ms={ms / max(n, 1):.2f} wall_s={wall:.3f} "
          f"samples/s={n_particles * n_trials * n_t / (ms * 1e-3):.3e} deterministic={np.array_equal(d[:5000], d2)} nan={int(np.isnan(d).sum())}", flush=True)
    np.save(f"gpurun_out/tc_check_{tag}.npy", d)


if __name__ == "__main__":
    if len(sys.argv) > 3:
        child(sys.argv[1], int(sys.argv[2]), sys.argv[3])
    else:
        n = sys.argv[1] if len(sys.argv) > 1 else "20000"
        os.makedirs("gpurun_out", exist_ok=True)
        runs = [("0", "ffma", {}), ("1", "tc", {}), ("1", "tc_seg4", {"ITJL_TC_SEG": "4"}), ("1", "tc_seg16", {"ITJL_TC_SEG": "16"}),
                ("1", "tc_seg52", {"ITJL_TC_SEG": "52"})]
        for mode, tag, env in runs:
            e = dict(os.environ); e.update(env)
            subprocess.run([sys.executable, __file__, mode, n, tag], check=False, timeout=900, env=e)
        try:
            a = np.load("gpurun_out/tc_check_ffma.npy")
            for _, tag, _ in runs[1:]:
                b = np.load(f"gpurun_out/tc_check_{tag}.npy")
                rel = np.abs(a - b) / np.maximum(np.abs(a), 1e-30)
                print(tag, "distance rel diff vs ffma: max", np.nanmax(rel), "median", np.nanmedian(rel), "nan mismatch", int((np.isnan(a) != np.isnan(b)).sum()))
        except Exception as e:  # noqa: BLE001
            print("compare failed:", e)
