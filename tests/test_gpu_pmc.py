"""GPU parity of the between-generation step: calc_weights (src/core/abc.jl:520-567),
weighted_covar (:582-613) on the device against the oracle's fp64 restatement."""
import numpy as np
import pytest

from oracle import abc as oabc
from oracle import models as omodels

pytestmark = pytest.mark.gpu


def _case(rng, n_prev, n, p):
    th_prev = np.stack([rng.gamma(4.0, 0.1, n_prev)] + [rng.normal(10.0, 2.0, n_prev) for _ in range(p - 1)], axis=1)
    th = np.stack([rng.gamma(4.0, 0.1, n)] + [rng.normal(10.0, 2.0, n) for _ in range(p - 1)], axis=1)
    w = rng.uniform(0.2, 1.8, n_prev); w /= w.sum()
    cov = 2.0 * np.atleast_2d(np.cov(th_prev, rowvar=False)) + 1e-6 * np.eye(p)
    return th_prev, th, w, cov


def _priors(pkg, p):
    po = [omodels.Uniform(0.0, 5.0), omodels.Normal(10.0, 5.0), omodels.Normal(9.0, 3.0), omodels.Uniform(0.0, 30.0)][:p]
    pp = [pkg.Uniform(0.0, 5.0), pkg.Normal(10.0, 5.0), pkg.Normal(9.0, 3.0), pkg.Uniform(0.0, 30.0)][:p]
    return po, pp


@pytest.mark.parametrize("n_prev,n,p", [(100, 100, 1), (60, 40, 3), (257, 1031, 2), (1500, 700, 4), (3000, 5000, 1)])
def test_calc_weights_fp64_path_matches_oracle(pkg, ctx, n_prev, n, p):
    rng = np.random.default_rng(n_prev + n + p)
    th_prev, th, w, cov = _case(rng, n_prev, n, p)
    po, pp = _priors(pkg, p)
    want = oabc.calc_weights(th_prev, th, cov, w, po)
    got = pkg.calc_weights(th_prev, th, cov, w, pp, ctx=ctx)
    assert got.shape == want.shape and np.isclose(got.sum(), 1.0, rtol=1e-12)
    assert np.allclose(got, want, rtol=1e-10, atol=0)
    # un-normalised: the Gaussian-mixture denominator itself (what a sharded caller gathers)
    raw = pkg.calc_weights(th_prev, th, cov, w, pp, ctx=ctx, normalize=False)
    assert np.allclose(raw / raw.sum(), want, rtol=1e-10)
    assert np.array_equal(pkg.calc_weights(th_prev, th, cov, w, pp, ctx=ctx), got)        # deterministic (no atomics)


@pytest.mark.parametrize("p", [1, 2])
def test_calc_weights_fp32_pair_sums_above_1e8_pairs(pkg, ctx, p):
    """n * n_prev > 1e8 switches the pair sums to fp32 (one MUFU.EX2 per pair): whitened O(1) coordinates keep the
    weights within 1e-5 relative of the fp64 oracle (measured ~1e-6); tuning.pmc_fp32 = 0 forces fp64."""
    rng = np.random.default_rng(7 + p)
    th_prev, th, w, cov = _case(rng, 12000, 9000, p)
    po, pp = _priors(pkg, p)
    idx = np.arange(0, 9000, 37)
    want_den = np.array([np.sum(w * oabc.mvnormal_pdf(th[i][None, :], th_prev, cov)) for i in idx])
    pri = np.ones(9000)
    for j, pr in enumerate(po):
        pri *= pr.pdf(th[:, j])
    raw = pkg.calc_weights(th_prev, th, cov, w, pp, ctx=ctx, normalize=False)
    rel = np.abs(raw[idx] * want_den / np.where(pri[idx] > 0, pri[idx], 1.0) - (pri[idx] > 0))
    print(f"fp32 pair sums, p={p}: max relative error of a weight {rel.max():.3g}")
    assert rel.max() < 1e-5
    with ctx.tuned(pmc_fp32=0):
        raw64 = pkg.calc_weights(th_prev, th, cov, w, pp, ctx=ctx, normalize=False)
    rel64 = np.abs(raw64[idx] * want_den / np.where(pri[idx] > 0, pri[idx], 1.0) - (pri[idx] > 0))
    assert rel64.max() < 1e-11
    got = pkg.calc_weights(th_prev, th, cov, w, pp, ctx=ctx)
    assert np.isclose(got.sum(), 1.0, rtol=1e-12) and np.allclose(got, raw64 / raw64.sum(), rtol=1e-5)


def test_calc_weights_one_parameter_grid_path(pkg, ctx):
    """One-parameter models above 4e9 pairs (C5: 4.4e11): the mixture density is evaluated exactly in fp64 on an
    8192-point grid of the whitened coordinate and read off by 4-point interpolation - tuning.pmc_fp32 = 2 forces that
    path here at a size the oracle can check.  Interpolation error ~1e-12 of the density's scale: every weight within
    1e-8 of the fp64 pair sums, including proposals far out in the tails and a duplicated particle."""
    rng = np.random.default_rng(21)
    th_prev, th, w, cov = _case(rng, 6000, 20000, 1)
    th[:40, 0] = th_prev[:, 0].mean() + np.linspace(-9, 9, 40) * th_prev[:, 0].std()      # tails (most have prior 0)
    th[41] = th[40]
    po, pp = _priors(pkg, 1)
    want = oabc.calc_weights(th_prev, th, cov, w, po)
    with ctx.tuned(pmc_fp32=0):
        exact = pkg.calc_weights(th_prev, th, cov, w, pp, ctx=ctx, normalize=False)
    with ctx.tuned(pmc_fp32=2):
        n0 = ctx.launch_count
        raw = pkg.calc_weights(th_prev, th, cov, w, pp, ctx=ctx, normalize=False)
        assert ctx.launch_count - n0 == 3                    # mixture on the grid, density, interpolation (one GPU)
        got = pkg.calc_weights(th_prev, th, cov, w, pp, ctx=ctx)
        again = pkg.calc_weights(th_prev, th, cov, w, pp, ctx=ctx)
    ok = exact > 0
    rel = np.abs(raw[ok] - exact[ok]) / exact[ok]
    print(f"grid path: max relative error of a weight {rel.max():.3g} over {ok.sum()} particles")
    assert rel.max() < 1e-8 and np.array_equal(raw == 0, exact == 0)
    assert np.isclose(got.sum(), 1.0, rtol=1e-12) and np.allclose(got, want, rtol=1e-8, atol=0)
    assert np.array_equal(got, again) and got[40] == got[41]
    # a degenerate new population (all particles equal) has no grid: the pair sums take over
    same = np.full((64, 1), th[100, 0])
    with ctx.tuned(pmc_fp32=2):
        d = pkg.calc_weights(th_prev, same, cov, w, pp, ctx=ctx)
    assert np.allclose(d, 1.0 / 64, rtol=1e-12)


def test_weighted_covar_and_reference_kats(pkg, ctx):
    rng = np.random.default_rng(3)
    x = rng.normal(size=(5000, 3)) * np.array([1.0, 5.0, 0.1]) + np.array([0.3, 10.0, -2.0])
    w = rng.uniform(0.1, 2.0, 5000)
    assert np.allclose(pkg.weighted_covar(x, w, ctx=ctx), oabc.weighted_covar(x, w), rtol=1e-11, atol=1e-14)
    assert np.allclose(pkg.weighted_covar(x[:, :1], w, ctx=ctx), oabc.weighted_covar(x[:, :1], w), rtol=1e-11)
    # equal weights = the unbiased sample covariance (test/test_abc.jl:66-75)
    assert np.allclose(pkg.weighted_covar(x, np.ones(5000), ctx=ctx), np.cov(x, rowvar=False), rtol=1e-10)
    # the weights of a real calc_weights call sum to 1 (test_abc.jl:77-89)
    th_prev, th, wp, cov = _case(rng, 80, 90, 2)
    _, pp = _priors(pkg, 2)
    wn = pkg.calc_weights(th_prev, th, cov, wp, pp, ctx=ctx)
    assert wn.sum() == pytest.approx(1.0) and np.all(wn >= 0)
    with pytest.raises(pkg.ItjlError):
        pkg.calc_weights(th_prev, th, -cov, wp, pp, ctx=ctx)            # not positive definite


def test_pmc_abc_end_to_end_uses_the_device_weights(pkg, ctx):
    """A real pmc_abc run (fused kernels + device weights) against the same run with the oracle's weight functions
    injected: identical accepted sets, weights within fp64 rounding."""
    from oracle import ou as oou
    dt, n_t = 0.002, 2000
    data = oou.generate_ou_process(0.1, 1.0, dt, n_t * dt, 4, seed=5, n_t=n_t)
    t = dt * np.arange(1, n_t + 1)
    model = pkg.one_timescale_model(data, t, "abc", prior=[pkg.Uniform(0.01, 1.0)])
    kw = dict(epsilon_0=1.0, max_iter=2000, min_accepted=60, steps=4, target_epsilon=1e-6, seed=3, ctx=ctx)
    a = pkg.pmc_abc(model, rng=np.random.default_rng(1), **kw)
    b = pkg.pmc_abc(model, rng=np.random.default_rng(1), calc_weights_fn=oabc.calc_weights,
                    weighted_covar_fn=oabc.weighted_covar, **kw)
    assert len(a.epsilon_history) == len(b.epsilon_history) >= 2
    assert np.array_equal(a.theta_history[1], b.theta_history[1])
    assert np.allclose(a.weights_history[1], b.weights_history[1], rtol=1e-9)


def test_device_pmc_proposals_have_the_reference_distribution(pkg, ctx):
    """itjl_pmc_propose against draw_theta_pmc's definition (abc.jl:101-117): resampling frequencies follow the weights,
    the perturbation has covariance tau_squared + jitter I, nothing negative comes out, and proposal i depends only on
    (seed, generation, first + i)."""
    rng = np.random.default_rng(0)
    n_prev, n = 500, 400000
    theta_prev = np.stack([rng.uniform(0.5, 3.0, n_prev), rng.uniform(5.0, 15.0, n_prev)], axis=1)
    w = rng.random(n_prev) ** 3
    w /= w.sum()
    tsq = np.array([[0.04, 0.03], [0.03, 0.25]])
    th = pkg.draw_theta_pmc_device(theta_prev, w, tsq, n, seed=11, generation=2, ctx=ctx)
    assert th.shape == (n, 2) and np.all(th >= 0) and np.all(np.isfinite(th))
    # batching invariance
    a = pkg.draw_theta_pmc_device(theta_prev, w, tsq, 1000, seed=11, generation=2, first=0, ctx=ctx)
    b = pkg.draw_theta_pmc_device(theta_prev, w, tsq, 600, seed=11, generation=2, first=400, ctx=ctx)
    assert np.array_equal(a, th[:1000]) and np.array_equal(b, th[400:1000])
    assert not np.array_equal(pkg.draw_theta_pmc_device(theta_prev, w, tsq, 1000, seed=11, generation=3, ctx=ctx), a)
    # moments of the mixture sum_j w_j N(theta_j, C) (nothing is near zero here, so no draw was rejected):
    # mean = sum w theta, covariance = C + weighted covariance of theta_prev
    mu = w @ theta_prev
    cov = tsq + 1e-5 * np.eye(2) + (theta_prev - mu).T @ ((theta_prev - mu) * w[:, None])
    se = np.sqrt(np.diag(cov) / n)
    assert np.all(np.abs(th.mean(axis=0) - mu) < 5 * se)
    assert np.allclose(np.cov(th, rowvar=False), cov, rtol=0.02)
    # resampling frequencies: nearest theta_prev after removing the perturbation is not identifiable, so test through a
    # degenerate covariance instead
    t0 = pkg.draw_theta_pmc_device(theta_prev, w, 1e-12 * np.eye(2), 200000, seed=5, generation=1, jitter=0.0, ctx=ctx)
    idx = np.argmin(np.abs(t0[:, :1] - theta_prev[:, 0][None, :]), axis=1)
    freq = np.bincount(idx, minlength=n_prev) / 200000
    assert np.max(np.abs(freq - w) / np.sqrt(w * (1 - w) / 200000 + 1e-12)) < 5.5
    # redraw-while-negative: theta* close to zero, wide proposal - the accepted draws are the positive part of the Gaussian
    t1 = pkg.draw_theta_pmc_device(np.array([[0.05]]), np.array([1.0]), np.array([[1.0]]), 100000, seed=3, ctx=ctx)
    assert t1.min() >= 0 and abs(t1.mean() - (0.05 + np.exp(-0.05 ** 2 / 2) / np.sqrt(2 * np.pi) / 0.5199388)) < 0.01
    with pytest.raises(pkg.ItjlError):
        pkg.draw_theta_pmc_device(theta_prev, w, -tsq, 10, ctx=ctx)


def test_pmc_abc_with_device_proposals(pkg, ctx):
    """pmc_abc(device_proposals = True) recovers the timescale like the host-proposal run does."""
    from oracle import ou as oou
    dt, n_t = 0.002, 2500
    t = dt * np.arange(1, n_t + 1)
    data = oou.generate_ou_process(0.1, 1.0, dt, n_t * dt, 20, seed=4, n_t=n_t)
    model = pkg.one_timescale_model(data, t, "abc", prior=[pkg.Uniform(0.01, 1.0)])
    res = [pkg.pmc_abc(model, epsilon_0=1.0, max_iter=4000, min_accepted=300, steps=6, rng=np.random.default_rng(1), seed=2,
                       ctx=ctx, device_proposals=dp) for dp in (False, True)]
    m = [float(np.average(r.final_theta[:, 0], weights=r.final_weights)) for r in res]
    assert abs(m[0] - 0.1) < 0.03 and abs(m[1] - 0.1) < 0.03 and abs(m[0] - m[1]) < 0.03


def test_select_epsilon_device_order_statistics_match_numpy(pkg, ctx):
    """select_epsilon(..., ctx): the percentiles of a large generation come from a device sort and are the same
    floats np.percentile (quantile type 7 = StatsBase.percentile) gives, NaN and >= distance_max removed first."""
    rng = np.random.default_rng(9)
    d = rng.gamma(2.0, 0.3, size=300001)
    d[rng.integers(0, d.size, 5000)] = np.nan
    d[rng.integers(0, d.size, 3000)] = 25.0                    # >= distance_max: dropped
    for it, rate, eps in [(1, 0.0, 1.0), (3, 0.5, 0.6), (3, 0.001, 0.05), (4, 0.0101, 0.3)]:
        want = pkg.select_epsilon(d, eps, current_acc_rate=rate, iteration=it, total_iterations=10)
        got = pkg.select_epsilon(d, eps, current_acc_rate=rate, iteration=it, total_iterations=10, ctx=ctx)
        assert got == want, (it, rate, got, want)
    allbad = np.full(70000, np.nan)
    assert pkg.select_epsilon(allbad, 0.7, iteration=2, current_acc_rate=0.5, ctx=ctx) == 0.7
