"""north_star (3): posterior timescale mean/std from the GPU path within Monte-Carlo error
of the CPU oracle running the reference's serial PMC-ABC (README-style OU example)."""
import numpy as np
import pytest

from oracle import abc as oabc
from oracle import cport
from oracle import models as omodels
from oracle import ou as oou

pytestmark = pytest.mark.gpu

DT = 0.002
PARAMS = dict(epsilon_0=1.0, max_iter=3000, min_accepted=100, steps=8, target_epsilon=1e-4,
              convergence_window=5, minAccRate=0.01)


def _fit_oracle(om, seed):
    rng = np.random.default_rng(1000 + seed)

    def batch_fn(thetas, first, generation):
        return cport.abc_distances(om, thetas, seed=seed, generation=generation, particle_offset=first)
    rec = oabc.pmc_abc(om.prior, None, rng, distance_batch_fn=batch_fn, **PARAMS)
    th = rec[-1]["theta_accepted"][:, 0]
    w = rec[-1]["weights"]
    return th, w


def _fit_gpu(pkg, ctx, gm, seed):
    res = pkg.pmc_abc(gm, rng=np.random.default_rng(2000 + seed), seed=seed, ctx=ctx, **PARAMS)
    return res.final_theta[:, 0], res.final_weights


def test_posterior_matches_oracle_within_mc_error(pkg, ctx):
    n_t, n_trials, tau = 2000, 4, 0.1
    t = DT * np.arange(1, n_t + 1)
    data = oou.generate_ou_process(tau, 1.0, DT, n_t * DT, n_trials, seed=123, n_t=n_t)
    om = omodels.one_timescale_model(data, t, prior=[omodels.Uniform(0.01, 1.0)])
    gm = pkg.one_timescale_model(data, t, "abc", prior=[pkg.Uniform(0.01, 1.0)])
    n_rep = 6
    mo, so, mg, sg = [], [], [], []
    for s in range(n_rep):
        th, w = _fit_oracle(om, s)
        mo.append(np.average(th, weights=w)); so.append(th.std())
        th, w = _fit_gpu(pkg, ctx, gm, s)
        mg.append(np.average(th, weights=w)); sg.append(th.std())
    mo, so, mg, sg = map(np.array, (mo, so, mg, sg))
    # both recover the generating timescale
    assert abs(mo.mean() - tau) < 0.05 and abs(mg.mean() - tau) < 0.05
    # posterior means agree within Monte-Carlo error (3 sigma of the difference of means)
    se = np.sqrt(mo.var(ddof=1) / n_rep + mg.var(ddof=1) / n_rep)
    assert abs(mo.mean() - mg.mean()) < 3 * se + 0.1 * so.mean()
    # posterior widths agree within a factor the replicate scatter supports
    assert 0.5 < sg.mean() / so.mean() < 2.0


def test_posterior_predictive_statistics(pkg, ctx):
    """posterior_predictive's numbers (IntPlottingExt.jl:105-128): batched summary statistics at posterior
    draws; mean and 95 % band bracket the data's ACF for a posterior concentrated at the true timescale."""
    import numpy as np
    from oracle import ou as oou
    dt, n_t, n_trials = 0.002, 5000, 10
    data = oou.generate_ou_process(0.3, 1.0, dt, n_t * dt, n_trials, seed=2, n_t=n_t)
    t = dt * np.arange(1, n_t + 1)
    model = pkg.one_timescale_model(data, t, "abc", prior=[pkg.Uniform(0.05, 5.0)])
    post = pkg.ABCResults([], [], [], [], np.random.default_rng(0).normal(0.3, 0.02, size=(500, 1)), np.ones(500) / 500, np.array([0.3]))
    pp = pkg.posterior_predictive(post, model, n_samples=200, rng=np.random.default_rng(1), seed=5, ctx=ctx)
    assert pp["predictions"].shape == (200, model.n_lags) and pp["theta_samples"].shape == (200, 1)
    assert np.all(pp["pred_lower"] <= pp["pred_mean"]) and np.all(pp["pred_mean"] <= pp["pred_upper"])
    assert np.allclose(pp["predictions"][:, 0], 1.0, atol=1e-5)
    inside = (model.data_sum_stats >= pp["pred_lower"] - 1e-9) & (model.data_sum_stats <= pp["pred_upper"] + 1e-9)
    assert inside.mean() > 0.5          # neighbouring lags are strongly correlated: a sanity check, not a coverage test
    # same numbers as scoring the same draws directly
    _, stats = model.gpu_model(ctx).distances(pp["theta_samples"], seed=5, generation=0, return_stats=True)
    assert np.array_equal(stats, pp["predictions"])
    # against the oracle: the reference's loop (generate_data -> summary_stats per draw, IntPlottingExt.jl:114-124) on the
    # same draws and Philox streams, for the ACF model and for the periodogram (PSD) model
    from oracle import models as omodels
    for summary_method in ("acf", "psd"):
        om = omodels.one_timescale_model(data, t, summary_method=summary_method, prior=[omodels.Uniform(0.05, 5.0)])
        gm = pkg.one_timescale_model(data, t, "abc", summary_method=summary_method, prior=[pkg.Uniform(0.05, 5.0)])
        pq = pkg.posterior_predictive(post, gm, n_samples=16, rng=np.random.default_rng(3), seed=8, ctx=ctx)
        want = np.stack([omodels.summary_stats(om, omodels.generate_data(om, th, seed=8, generation=0, particle=i))
                         for i, th in enumerate(pq["theta_samples"])])
        scale = 1.0 if summary_method == "acf" else want.max()
        assert np.max(np.abs(pq["predictions"] - want)) <= 1e-5 * scale
        assert np.max(np.abs(pq["pred_mean"] - want.mean(axis=0))) <= 1e-5 * scale
        assert np.max(np.abs(pq["pred_lower"] - np.quantile(want, 0.025, axis=0))) <= 1e-5 * scale
        assert np.max(np.abs(pq["pred_upper"] - np.quantile(want, 0.975, axis=0))) <= 1e-5 * scale
