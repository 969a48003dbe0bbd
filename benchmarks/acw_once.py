#!/usr/bin/env python
"""One acw() call on randn(10000 x 5000) - profiling target for the K4 kernels."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge
pkg = ge.load_package()
ctx = pkg.default_context(0)
d = np.asfortranarray(np.random.default_rng(0).standard_normal((10000, 5000)))
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 2):
    r = pkg.acw(d, 100.0, acwtypes=["acw0", "acw50", "acweuler", "tau"], ctx=ctx)
print("n_lags", r.n_lags)
