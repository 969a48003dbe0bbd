import os, sys, numpy as np
sys.path.insert(0, os.getcwd())
import __graft_entry__ as ge
from oracle import ou as oou
pkg = ge.load_package(); ctx = pkg.default_context(0)
dt, n_t, n_trials = 1 / 500, 5000, 20
t = dt * np.arange(1, n_t + 1)
data = oou.generate_ou_with_oscillation([0.05, 10.0, 0.7], dt, n_t * dt, n_trials, 0.0, 1.0, seed=23, n_t=n_t)
model = pkg.one_timescale_and_osc_model(data, t, "abc", summary_method="psd", ctx=ctx)
pri = [pkg.Uniform(0.005, 0.3), model.prior[1], pkg.Uniform(0.0, 1.0)]
m2 = pkg.one_timescale_and_osc_model(data, t, "abc", summary_method="psd", prior=pri, ctx=ctx)
for steps in (8, 14, 20):
    res = pkg.int_fit(m2, dict(max_iter=6000, min_accepted=300, steps=steps, show_progress=False, verbose=False, target_epsilon=1e-4),
                      rng=np.random.default_rng(4), seed=9, ctx=ctx)
    w = res.final_weights
    print(steps, [float(np.average(res.final_theta[:, c], weights=w)) for c in range(3)], [round(e, 4) for e in res.epsilon_history])
