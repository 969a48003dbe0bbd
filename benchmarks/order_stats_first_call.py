#!/usr/bin/env python
"""Diagnostic: time of the FIRST itjl_order_stats call of a process (after a 10^6-particle generation), split in two."""
import os, sys, time, ctypes
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge
pkg = ge.load_package(); ctx = pkg.default_context(0)
rng = np.random.default_rng(1)
d = np.abs(rng.standard_normal(1_000_000))
d[::50] = np.nan
d = np.ascontiguousarray(d, dtype=np.float64)
L = pkg._lib
for rep in range(3):
    t0 = time.perf_counter()
    probs = np.array([0.25, 0.5, 0.75]); pairs = np.empty(6); nv = ctypes.c_int64(0)
    ctx.check(L.lib().itjl_order_stats(ctx.handle, L.ptr(d), d.size, 10.0, L.ptr(probs), 3, L.ptr(pairs), ctypes.byref(nv)))
    t1 = time.perf_counter()
    e = pkg.select_epsilon(d, 1.0, ctx=ctx)
    t2 = time.perf_counter()
    print(f"rep {rep}: itjl_order_stats {1e3 * (t1 - t0):.1f} ms, select_epsilon {1e3 * (t2 - t1):.1f} ms", flush=True)
