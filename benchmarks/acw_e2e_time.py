#!/usr/bin/env python
"""C2 acw() from a host array (400 MB host -> device inside the call): ms per call, effective upload rate."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge
pkg = ge.load_package()
ctx = pkg.default_context(0)
d = np.asfortranarray(np.random.default_rng(7).standard_normal((10000, 5000)))
types = ["acw0", "acw50", "acweuler", "auc", "tau"]
for threads in (0, 4, 8, 16):
    ctx.set_tuning(copy_threads=threads)
    pkg.acw(d, 100.0, acwtypes=types, ctx=ctx)
    t0 = time.perf_counter()
    for _ in range(5):
        r = pkg.acw(d, 100.0, acwtypes=types, ctx=ctx)
    s = (time.perf_counter() - t0) / 5
    print("copy_threads", threads, "ms per call", round(1e3 * s, 2), "GB/s", round(0.4 / s, 1), "slices/s", round(10000 / s), flush=True)
