#!/usr/bin/env python
"""Tensor-core vs FFMA fused ACF kernel on a C5-shaped model: per-lag agreement and timing."""
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def child(mode, n_particles):
    os.environ["ITJL_ABC_TC"] = mode
    import __graft_entry__ as ge
    import bench
    pkg = ge.load_package()
    ctx = pkg.default_context(0)
    data, t = bench.observed_data_host()
    model = pkg.one_timescale_model(data, t, "abc", prior=[pkg.Uniform(0.05, 5.0)])
    gm = model.gpu_model(ctx)
    rng = np.random.default_rng(3)
    theta = rng.uniform(0.05, 5.0, size=(n_particles, 1))
    d8, st8 = gm.distances(theta[:8], seed=9, generation=1, return_stats=True)
    gm.distances(theta[:296], seed=9, generation=1)
    ctx.timing_reset()
    t0 = time.perf_counter()
    d = gm.distances(theta, seed=9, generation=1)
    wall = time.perf_counter() - t0
    n, ms = ctx.timing_get()
    np.save(f"gpurun_out/tc_check_{mode}.npy", np.concatenate([d8, st8.ravel(), d]))
    print(f"mode={mode} n_lags={model.n_lags} particles={n_particles} kernel_ms={ms / max(n, 1):.2f} wall_s={wall:.3f} "
          f"samples/s={n_particles * 50 * 5000 / (ms * 1e-3):.3e}", flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 2:
        child(sys.argv[1], int(sys.argv[2]))
    else:
        n = sys.argv[1] if len(sys.argv) > 1 else "20000"
        os.makedirs("gpurun_out", exist_ok=True)
        for mode in ("0", "1"):
            subprocess.run([sys.executable, __file__, mode, n], check=False, timeout=600)
        try:
            a = np.load("gpurun_out/tc_check_0.npy"); b = np.load("gpurun_out/tc_check_1.npy")
            nl = (len(a) - 8 - int(n)) // 8
            print("dist8 ffma", a[:8]); print("dist8 tc  ", b[:8])
            sa = a[8:8 + 8 * nl].reshape(8, nl); sb = b[8:8 + 8 * nl].reshape(8, nl)
            err = np.abs(sa - sb)
            print("max |acf_tc - acf_ffma| per particle:", err.max(axis=1))
            print("argmax lag:", err.argmax(axis=1), " mean signed diff:", (sb - sa).mean())
            da, db = a[8 + 8 * nl:], b[8 + 8 * nl:]
            rel = np.abs(da - db) / np.maximum(np.abs(da), 1e-30)
            print("all-particle distance rel diff: max", np.nanmax(rel), "median", np.nanmedian(rel), "nan mismatch", int((np.isnan(da) != np.isnan(db)).sum()))
        except Exception as e:  # noqa: BLE001
            print("compare failed:", e)
