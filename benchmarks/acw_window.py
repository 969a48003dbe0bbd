import os, sys, time
import numpy as np
sys.path.insert(0, '/root/repo')
import __graft_entry__ as ge
pkg = ge.load_package()
ctx = pkg.default_context(0)
d = np.asfortranarray(np.random.default_rng(7).standard_normal((10000, 5000)))
types = ["acw0", "acw50", "acweuler", "auc", "tau"]
dd = pkg.upload(d, ctx=ctx)
base = None
for w in (1, 32, 16, 8):
    ctx.set_tuning(acw_stream=w)
    r = pkg.acw(dd, 100.0, acwtypes=types)
    ctx.timing_reset()
    t0 = time.perf_counter()
    for _ in range(10):
        r = pkg.acw(dd, 100.0, acwtypes=types)
    s = (time.perf_counter() - t0) / 10
    n, ms = ctx.timing_get(); ctx.timing_stop()
    res = [np.asarray(x) for x in r.acw_results]
    if base is None: base = res
    same = all(np.array_equal(a, b, equal_nan=True) for a, b in zip(res[:3], base[:3])) and all(np.allclose(a, b, rtol=1e-9, equal_nan=True) for a, b in zip(res[3:], base[3:]))
    print("W", w, "ms/call", round(1e3 * s, 3), "stream+finish us", round(1e3 * ms / max(n, 1), 1), "frac", round(4e8 / (ms / max(n, 1) * 1e-3) / 6453e9, 3), "n_lags", r.n_lags, "same", same, flush=True)
