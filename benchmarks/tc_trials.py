#!/usr/bin/env python
"""Tensor-core vs CUDA-core fused ACF kernel as a function of the number of trials (diagnostic)."""
import os, subprocess, sys
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


def child(n_trials, n_lags, n_particles):
    import __graft_entry__ as ge
    from oracle import ou as oou
    pkg = ge.load_package()
    ctx = pkg.default_context(0)
    dt, n_t = 0.002, 5000
    data = oou.generate_ou_process(0.3, 1.0, dt, n_t * dt, n_trials, seed=3, n_t=n_t)
    m = pkg.one_timescale_model(data, dt * np.arange(1, n_t + 1), "abc", prior=[pkg.Uniform(0.05, 5.0)], n_lags=n_lags)
    g = m.gpu_model(_tun(ctx))
    theta = np.random.default_rng(0).uniform(0.05, 5.0, size=(n_particles, 1))
    g.distances(theta[:296], seed=1, generation=1)
    ctx.timing_reset()
    g.distances(theta, seed=1, generation=1)
    n, ms = ctx.timing_get()
    print(f"TC={os.environ['ITJL_ABC_TC']} kernel={g.kernel_info()[0]} n_trials={n_trials} n_lags={n_lags} ms={ms / n:.2f} "
          f"samples/s={n_particles * n_trials * n_t / (ms / n * 1e-3):.3e}", flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "child":
        child(int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]))
    else:
        for nt in (1, 2, 3, 4, 6, 8, 16):
            for tc in ("0", "1"):
                e = dict(os.environ); e["ITJL_ABC_TC"] = tc
                subprocess.run([sys.executable, __file__, "child", str(nt), "400", str(max(2960, 59200 // nt))], env=e, timeout=300)
