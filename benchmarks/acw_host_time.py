import os, sys, time, ctypes
import numpy as np
sys.path.insert(0, '/root/repo')
import __graft_entry__ as ge
pkg = ge.load_package()
ctx = pkg.default_context(0)
d = np.asfortranarray(np.random.default_rng(7).standard_normal((10000, 5000)))
types = ["acw0", "acw50", "acweuler", "auc", "tau"]
dd = pkg.upload(d, ctx=ctx)
L = pkg._lib.lib()
orig = L.itjl_acw_resident
acc = [0.0, 0]
def timed(*a):
    t0 = time.perf_counter(); r = orig(*a); acc[0] += time.perf_counter() - t0; acc[1] += 1; return r
class Proxy:
    def __getattr__(self, k):
        return timed if k == "itjl_acw_resident" else getattr(L, k)
pkg._lib.lib = lambda: Proxy()
for rep in range(2):
    acc[0] = 0; acc[1] = 0
    t0 = time.perf_counter()
    for _ in range(20):
        r = pkg.acw(dd, 100.0, acwtypes=types)
    tot = (time.perf_counter() - t0) / 20
    print("total ms/call", 1e3 * tot, "inside C call ms", 1e3 * acc[0] / acc[1])
for kw in (dict(return_acf=False), dict(acwtypes=["acw0", "acw50"]), dict(acwtypes=["acw0", "acw50"], return_acf=False), dict(acwtypes=["tau"], return_acf=False)):
    acc[0] = 0; acc[1] = 0
    k = dict(acwtypes=types); k.update(kw)
    t0 = time.perf_counter()
    for _ in range(20):
        r = pkg.acw(dd, 100.0, **k)
    tot = (time.perf_counter() - t0) / 20
    print(kw, "total ms/call", round(1e3 * tot, 3), "inside C call ms", round(1e3 * acc[0] / acc[1], 3))
