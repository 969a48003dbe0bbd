#!/usr/bin/env python
"""ACF error of the tensor-core kernel against the CUDA-core kernel over many particles, as a function of
the accumulator segment length (diagnostic for the drain interval)."""
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


def child(tag):
    import __graft_entry__ as ge
    import bench
    pkg = ge.load_package()
    ctx = pkg.default_context(0)
    data, t = bench.observed_data_host()
    model = pkg.one_timescale_model(data, t, "abc", prior=[pkg.Uniform(0.05, 5.0)])
    g = model.gpu_model(_tun(ctx))
    theta = np.random.default_rng(7).uniform(0.05, 5.0, size=(1184, 1))
    d, s = g.distances(theta, seed=21, generation=4, return_stats=True)
    np.save(f"gpurun_out/tc_err_{tag}.npy", s)


if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[1] == "child":
        child(sys.argv[2])
    else:
        os.makedirs("gpurun_out", exist_ok=True)
        segs = [int(a) for a in sys.argv[1:]] or [8, 13, 17, 25, 50]
        runs = [("ffma", {"ITJL_ABC_TC": "0"})] + [(f"seg{k}", {"ITJL_ABC_TC": "1", "ITJL_TC_SEG": str(k)}) for k in segs]
        for tag, env in runs:
            e = dict(os.environ); e.update(env)
            subprocess.run([sys.executable, __file__, "child", tag], env=e, timeout=300)
        ref = np.load("gpurun_out/tc_err_ffma.npy")
        for tag, _ in runs[1:]:
            err = np.load(f"gpurun_out/tc_err_{tag}.npy") - ref
            slope = float((ref * err).sum() / (ref * ref).sum())        # bias proportional to the value
            print(f"{tag}: max |err| {np.abs(err).max():.2e}  rms {np.sqrt(np.mean(err ** 2)):.2e}  mean {err.mean():+.2e}  "
                  f"99.9% {np.quantile(np.abs(err), 0.999):.2e}  slope {slope:+.2e}  "
                  + "  ".join(f"ac~{lo:.1f}: {np.mean(err[(ref >= lo) & (ref < hi)]):+.2e}" for lo, hi in ((0.9, 1.01), (0.5, 0.9), (0.1, 0.5))))
