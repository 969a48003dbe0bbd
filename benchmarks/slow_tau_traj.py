#!/usr/bin/env python
"""Near-unit-root trials (tau >> T): where does the fp32 path lose accuracy?  GPU generator (fp32 blocked scan) vs the
fp64 oracle trajectory, with injected fp64 noise and with device Philox noise; ACF of both trajectories in fp64."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge
from oracle import ou as oou, summary as osum
pkg = ge.load_package()
dt, n_t = 0.002, 5000
for tau in (3.0, 20.0, 500.0):
    for seed in range(3):
        u0, _, noise = oou.draw_inputs(seed, 0, 0, 1, n_t)
        for mode in ("injected", "philox"):
            kw = dict(u0=u0, noise=noise) if mode == "injected" else dict(seed=seed)
            xo = oou.generate_ou_process(tau, 1.0, dt, n_t * dt, 1, u0=u0, noise=noise, n_t=n_t)
            xg = pkg.generate_ou_process(tau, 1.0, dt, n_t * dt, 1, **kw)
            ao, ag = osum.comp_ac_fft(xo, 415), osum.comp_ac_fft(xg, 415)
            xr = oou.generate_ou_process(tau, 1.0, dt, n_t * dt, 1, standardize=False, u0=u0, noise=noise, n_t=n_t)
            gr = pkg.generate_ou_process(tau, 1.0, dt, n_t * dt, 1, standardize=False, **kw)
            print(f"tau={tau:6.1f} seed={seed} {mode:8s}: raw traj max|d| {np.abs(gr - xr).max():.2e} (range {np.ptp(xr):.2f})  "
                  f"std traj max|d| {np.abs(xg - xo).max():.2e}  ACF(fp64 of GPU traj) max|d| {np.abs(ag - ao).max():.2e} at lag {int(np.abs(ag - ao).argmax())}")
