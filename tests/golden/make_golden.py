#!/usr/bin/env python
"""Generates tests/golden/oracle_v1.npz: small frozen input/output vectors of the CPU oracle.

The reference is Julia and cannot be imported or run in this image, so these vectors are
regression fixtures of the oracle (itself pinned on the reference's KATs in
tests/test_oracle_reference_kats.py), not outputs of the reference.  Run from the repo root:
    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import acw, models, ou, philox, summary  # noqa: E402

dt, n_t = 0.002, 600
t = dt * np.arange(1, n_t + 1)
out = {}
out["philox_block"] = np.array(philox.philox4x32(np.arange(4), 1, 2, philox.c3_word(0, 3), 7, 9)).astype(np.uint32)
out["normals"] = philox.noise_normals(seed=7, generation=3, particle=2, trial=1, n=16)
u0, ph = philox.init_state_and_phase(7, 3, 2, 1)
out["u0_phase"] = np.array([u0, ph])
x = ou.generate_ou_process(0.05, 1.0, dt, n_t * dt, 3, seed=7, n_t=n_t)
out["ou"] = x
out["ou_em_raw"] = ou.generate_ou_process(0.05, 1.0, dt, n_t * dt, 2, standardize=False, seed=8, update="em", n_t=n_t)
out["ou_osc"] = ou.generate_ou_with_oscillation([0.05, 10.0, 0.8], dt, n_t * dt, 2, 0.5, 2.0, seed=9, n_t=n_t)
out["acf"] = summary.comp_ac_fft(x, 50)
xm = x.copy(); xm[0, 100:160] = np.nan; xm[2, 5] = np.nan
out["data_missing"] = xm
out["acf_missing"] = summary.comp_ac_time_missing(xm, 50)
p, f = summary.comp_psd_adfriendly(x, 1 / dt)
out["psd"], out["psd_freqs"] = p, f
m = models.one_timescale_model(x, t, prior=[models.Uniform(0.01, 1.0)])
out["model_n_lags"] = np.array([m.n_lags])
out["model_stats"] = m.data_sum_stats
theta = np.array([[0.05], [0.2], [0.01]])
out["theta"] = theta
out["dist"] = np.array([models.generate_data_and_reduce(m, th, seed=3, generation=1, particle=i) for i, th in enumerate(theta)])
mm = models.one_timescale_with_missing_model(xm, t, prior=[models.Uniform(0.01, 1.0)])
out["dist_missing"] = np.array([models.generate_data_and_reduce(mm, th, seed=3, generation=1, particle=i) for i, th in enumerate(theta)])
r = acw.acw(x, 1 / dt)
out["acw_k"] = np.stack([r["k0"], r["k50"], r["ke"]])
out["acw_auc"], out["acw_tau"], out["acw_n_lags"] = r["auc"], r["tau"], np.array([r["n_lags"]])
np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "oracle_v1.npz"), **out)
print({k: v.shape for k, v in out.items()})
