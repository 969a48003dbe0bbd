import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge
from oracle import ou as oou
pkg = ge.load_package(); ctx = pkg.default_context(0)
n_s, n_t, fs = 512, 5000, 100.0
d = oou.generate_ou_process(0.3, 1.0, 1 / fs, n_t / fs, n_s, seed=3, n_t=n_t)
rng = np.random.default_rng(0)
for r in range(n_s):
    a = int(rng.integers(0, n_t - 500)); d[r, a:a + 500] = np.nan
t = (1 / fs) * np.arange(1, n_t + 1)
p, f = pkg.comp_psd_lombscargle(t, d, np.isnan(d), 1 / fs, ctx=ctx)
print(p.shape)
