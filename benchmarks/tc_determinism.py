#!/usr/bin/env python
"""Run-to-run and batch-size determinism of the tensor-core kernel (diagnostic)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["ITJL_ABC_TC"] = "1"
import __graft_entry__ as ge
import bench
pkg = ge.load_package()


def _tun(ctx=None):
    """ITJL_* environment switches of this script -> the context's itjl_tuning (benchmarks/_tuning.py)."""
    import os as _os, sys as _sys
    _sys.path.insert(0, _os.path.dirname(_os.path.abspath(__file__)))
    import _tuning
    return _tuning.from_env(ctx or pkg.default_context())

ctx = pkg.default_context(0)
data, t = bench.observed_data_host()
model = pkg.one_timescale_model(data, t, "abc", prior=[pkg.Uniform(0.05, 5.0)])
g = model.gpu_model(_tun(ctx))
rng = np.random.default_rng(7)
theta = rng.uniform(0.05, 5.0, size=(1184, 1))
a = g.distances(theta, seed=21, generation=4)
b = g.distances(theta, seed=21, generation=4)
print("same call twice: equal", np.array_equal(a, b), "n diff", int((a != b).sum()))
for n in (300, 148, 149, 296, 500, 1183):
    c = g.distances(theta[:n], seed=21, generation=4)
    d = a[:n] != c
    print(f"first {n}: n diff {int(d.sum())}", "idx", np.nonzero(d)[0][:10], "max rel", float(np.max(np.abs(a[:n] - c) / a[:n])))
a2, st = g.distances(theta, seed=21, generation=4, return_stats=True)
print("with stats vs without: n diff", int((a != a2).sum()), "max rel", float(np.max(np.abs(a - a2) / a)))
c = g.distances(theta[:300], seed=21, generation=4)
print("after stats call, first 300: n diff", int((a[:300] != c).sum()))
from oracle import ou as oou
data2 = oou.generate_ou_process(0.3, 1.0, 0.002, 10.0, 50, seed=3, n_t=5000)
m2 = pkg.one_timescale_model(data2, t, "abc", prior=[pkg.Uniform(0.05, 5.0)], n_lags=414)
g2 = m2.gpu_model(_tun(ctx))
x1, s1 = g2.distances(theta, seed=21, generation=4, return_stats=True)
x2 = g2.distances(theta[:300], seed=21, generation=4)
d = x1[:300] != x2
print("test model: n diff", int(d.sum()), np.nonzero(d)[0][:10], "max rel", float(np.max(np.abs(x1[:300] - x2) / x1[:300])))
x3 = g2.distances(theta, seed=21, generation=4)
print("test model full again: n diff", int((x1 != x3).sum()))
