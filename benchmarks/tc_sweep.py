#!/usr/bin/env python
"""Timing sweep of the tensor-core fused ACF kernel over n_lags (tensor-memory plan A / B, FFMA tail)
and the accumulator segment length (diagnostic)."""
import os
import subprocess
import sys

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


def child(n_lags, n_particles):
    import __graft_entry__ as ge
    import bench
    pkg = ge.load_package()
    ctx = pkg.default_context(0)
    data, t = bench.observed_data_host()
    model = pkg.one_timescale_model(data, t, "abc", prior=[pkg.Uniform(0.05, 5.0)], n_lags=n_lags)
    gm = model.gpu_model(_tun(ctx))
    rng = np.random.default_rng(3)
    theta = rng.uniform(0.05, 5.0, size=(n_particles, 1))
    gm.distances(theta[:296], seed=9, generation=1)
    ctx.timing_reset()
    gm.distances(theta, seed=9, generation=1)
    n, ms = ctx.timing_get()
    print(f"TC={os.environ.get('ITJL_ABC_TC')} SEG={os.environ.get('ITJL_TC_SEG', '-')} n_lags={model.n_lags} "
          f"kernel_ms={ms / max(n, 1):.2f} us/particle/SM={ms / max(n, 1) * 1e3 * 148 / n_particles:.1f} "
          f"samples/s={n_particles * data.shape[0] * data.shape[1] / (ms * 1e-3):.3e}", flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[1] == "child":
        child(int(sys.argv[2]), int(sys.argv[3]))
    else:
        n = sys.argv[1] if len(sys.argv) > 1 else "5920"
        lags = [int(x) for x in (sys.argv[2].split(",") if len(sys.argv) > 2 else "128,256,385,386,414,449,470".split(","))]
        segs = sys.argv[3].split(",") if len(sys.argv) > 3 else ["8", "52"]
        for L in lags:
            for seg in segs:
                e = dict(os.environ); e["ITJL_ABC_TC"] = "1"; e["ITJL_TC_SEG"] = seg
                subprocess.run([sys.executable, __file__, "child", str(L), n], check=False, timeout=600, env=e)
