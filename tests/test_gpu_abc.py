"""GPU parity: the fused aABC kernels (simulate -> ACF / NaN-aware ACF / PSD -> distance ->
accept) against the fp64 oracle, through the C ABI."""
import numpy as np
import pytest

from oracle import abc as oabc
from oracle import cport
from oracle import models as omodels
from oracle import ou as oou

pytestmark = pytest.mark.gpu

DT = 0.002
ACF_TOL = 1e-5        # north_star (1): ACF within 1e-5 (values in [-1, 1])


def dist_tol(d_ref, stat_tol=ACF_TOL):
    """|delta d| implied by a per-statistic error <= stat_tol for d = mean(e^2):
    2 sqrt(d) tol + tol^2, plus 1e-5 relative (north_star (2))."""
    d = np.maximum(np.nan_to_num(d_ref), 0.0)
    return 2.0 * np.sqrt(d) * stat_tol + stat_tol ** 2 + 1e-5 * d


def _time(n):
    return DT * np.arange(1, n + 1)


def make_plain(pkg, n_t=5000, n_trials=2, tau=0.3, seed=123, n_lags=None, update="exact"):
    data = oou.generate_ou_process(tau, 1.0, DT, n_t * DT, n_trials, seed=seed, n_t=n_t)
    om = omodels.one_timescale_model(data, _time(n_t), prior=[omodels.Uniform(0.05, 5)], n_lags=n_lags,
                                     update=update)
    gm = pkg.one_timescale_model(data, _time(n_t), "abc", prior=[pkg.Uniform(0.05, 5)], n_lags=n_lags,
                                 update=update)
    return om, gm


def test_model_constructor_matches_oracle(pkg, ctx):
    om, gm = make_plain(pkg)
    assert gm.n_lags == om.n_lags and gm.numTrials == om.numTrials
    assert np.max(np.abs(gm.data_sum_stats - om.data_sum_stats)) < 1e-12
    assert gm.data_sd == pytest.approx(om.data_sd) and gm.dt == pytest.approx(DT)
    data = om.data
    oi = omodels.one_timescale_model(data, om.time)           # informed prior via fit_expdecay
    gi = pkg.one_timescale_model(data, om.time, "abc")
    assert gi.prior[0].mu == pytest.approx(oi.prior[0].mu, rel=1e-9) and gi.prior[0].sigma == 20.0


@pytest.mark.parametrize("n_t,n_trials,n_lags", [(5000, 2, None), (1000, 5, 37), (2500, 1, 700), (601, 3, 20)])
def test_distances_with_injected_noise(pkg, ctx, n_t, n_trials, n_lags):
    om, gm = make_plain(pkg, n_t=n_t, n_trials=n_trials, n_lags=n_lags)
    rng = np.random.default_rng(n_t)
    theta = np.array([[0.3], [0.05], [1.5], [4.0], [0.011]])
    P = theta.shape[0]
    u0 = rng.standard_normal((P, n_trials))
    noise = rng.standard_normal((P, n_trials, n_t))
    want = cport.abc_distances(om, theta, u0=u0, noise=noise)
    got, stats = gm.gpu_model(ctx).distances(theta, u0=u0, noise=noise, return_stats=True)
    assert np.all(np.abs(got - want) <= dist_tol(want)), (got, want)
    # summary statistics themselves (mean ACF over trials)
    from oracle import summary as osum
    for i in range(P):
        x = oou.generate_ou_process(theta[i, 0], om.data_sd, DT, n_t * DT, n_trials, u0=u0[i],
                                    noise=noise[i], n_t=n_t)
        ac = osum.comp_ac_fft(x, om.n_lags).mean(axis=0)
        assert np.max(np.abs(stats[i] - ac)) <= ACF_TOL


@pytest.mark.parametrize("update", ["exact", "em"])
def test_distances_philox_mode_and_batching_invariance(pkg, ctx, update):
    om, gm = make_plain(pkg, n_t=3000, n_trials=3, update=update)
    rng = np.random.default_rng(1)
    theta = rng.uniform(0.05, 3.0, size=(48, 1))
    want = cport.abc_distances(om, theta, seed=42, generation=3)
    g = gm.gpu_model(ctx)
    got = g.distances(theta, seed=42, generation=3)
    assert np.all(np.abs(got - want) <= dist_tol(want))
    # results depend only on the GLOBAL particle index, not on batching / offsets
    a = g.distances(theta[:20], seed=42, generation=3, particle_offset=0)
    b = g.distances(theta[20:], seed=42, generation=3, particle_offset=20)
    assert np.array_equal(np.concatenate([a, b]), got)
    assert not np.array_equal(g.distances(theta, seed=43, generation=3), got)
    assert not np.array_equal(g.distances(theta, seed=42, generation=4), got)
    assert np.array_equal(g.distances(theta, seed=42, generation=3), got)        # deterministic


def test_negative_and_degenerate_tau(pkg, ctx):
    om, gm = make_plain(pkg, n_t=2000, n_trials=2)
    theta = np.array([[-0.5], [-2.0], [-1e-4], [1e-6], [50.0]])
    want = cport.abc_distances(om, theta, seed=3, generation=1)
    got = gm.gpu_model(ctx).distances(theta, seed=3, generation=1)
    assert np.isnan(got[2]) and np.isnan(want[2])             # overflow -> NaN -> never accepted
    ok = ~np.isnan(want)
    assert np.array_equal(np.isnan(got), np.isnan(want))
    assert np.all(np.abs(got[ok] - want[ok]) <= dist_tol(want[ok]) * 5)


def test_missing_data_model(pkg, ctx):
    n_t, n_trials = 3000, 6
    data = oou.generate_ou_process(0.3, 1.0, DT, n_t * DT, n_trials, seed=7, n_t=n_t)
    rng = np.random.default_rng(0)
    for r in range(n_trials):
        s = rng.integers(0, n_t - 300)
        data[r, s:s + 300] = np.nan                             # 10 % contiguous gap
    data[1, rng.integers(0, n_t, 50)] = np.nan
    om = omodels.one_timescale_with_missing_model(data, _time(n_t), prior=[omodels.Uniform(0.05, 5)])
    gm = pkg.one_timescale_with_missing_model(data, _time(n_t), "abc", prior=[pkg.Uniform(0.05, 5)])
    assert gm.n_lags == om.n_lags
    assert np.max(np.abs(gm.data_sum_stats - om.data_sum_stats)) < 1e-12
    assert gm.data_sd == pytest.approx(om.data_sd)
    theta = np.array([[0.3], [0.1], [2.0], [0.02]])
    P = theta.shape[0]
    u0 = rng.standard_normal((P, n_trials)); noise = rng.standard_normal((P, n_trials, n_t))
    want = cport.abc_distances(om, theta, u0=u0, noise=noise)
    got = gm.gpu_model(ctx).distances(theta, u0=u0, noise=noise)
    assert np.all(np.abs(got - want) <= dist_tol(want))
    want = cport.abc_distances(om, theta, seed=8, generation=2)
    got = gm.gpu_model(ctx).distances(theta, seed=8, generation=2)
    assert np.all(np.abs(got - want) <= dist_tol(want))


@pytest.mark.parametrize("summary_method", ["psd", "acf"])
def test_oscillation_model(pkg, ctx, summary_method):
    dt, n_t, n_trials = 1 / 500, 10000, 4
    t = dt * np.arange(1, n_t + 1)
    data = oou.generate_ou_with_oscillation([0.05, 10.0, 0.95], dt, n_t * dt, n_trials, 0.0, 1.0, seed=11, n_t=n_t)
    po = [omodels.Normal(.1, .1), omodels.Normal(10, 5), omodels.Uniform(0, 1)]
    pp = [pkg.Normal(.1, .1), pkg.Normal(10, 5), pkg.Uniform(0, 1)]
    om = omodels.one_timescale_and_osc_model(data, t, prior=po, summary_method=summary_method)
    gm = pkg.one_timescale_and_osc_model(data, t, "abc", prior=pp, summary_method=summary_method)
    assert om.data_sum_stats.shape == gm.data_sum_stats.shape
    if summary_method == "psd":
        assert np.array_equal(om.freq_idx, gm.freq_idx)
        rel = np.abs(gm.data_sum_stats - om.data_sum_stats) / om.data_sum_stats
        assert np.median(rel) < 1e-5
    theta = np.array([[0.05, 10.0, 0.95], [0.1, 12.0, 0.5], [0.02, 8.0, 1.2], [0.03, 9.5, -0.2], [0.5, 30.0, 0.99]])
    want = cport.abc_distances(om, theta, seed=5, generation=3)
    got, stats = gm.gpu_model(ctx).distances(theta, seed=5, generation=3, return_stats=True)
    if summary_method == "psd":
        # log-distance of PSDs: per-bin relative error ~1e-5 (fp32 FFT), d = mean(log-ratio^2)
        assert np.all(np.abs(got - want) <= dist_tol(want, 3e-5) * 2), (got, want)
    else:
        assert np.all(np.abs(got - want) <= dist_tol(want)), (got, want)


def test_accept_and_sequential_stop(pkg, ctx):
    om, gm = make_plain(pkg, n_t=1500, n_trials=2)
    rng = np.random.default_rng(5)
    theta = rng.uniform(0.05, 3.0, size=(700, 1))
    g = gm.gpu_model(ctx)
    d_all = g.distances(theta, seed=1, generation=2)
    for eps, need in [(np.median(d_all), 25), (np.min(d_all) * 0.5, 10), (np.max(d_all) * 2, 700), (np.median(d_all), 0),
                      (np.median(d_all), 351)]:
        d, isacc, n_total, n_acc = g.generation(theta, eps, need, seed=1, generation=2)
        assert np.array_equal(d, d_all)
        want_total = oabc.stop_index(d_all, eps, need) if need > 0 else theta.shape[0]
        assert n_total == want_total
        assert np.array_equal(isacc[:n_total], (d_all <= eps)[:n_total])
        assert n_acc == int((d_all <= eps)[:n_total].sum())
    # NaN distances are never accepted
    dn = d_all.copy(); dn[::3] = np.nan
    import ctypes
    isacc = np.empty(dn.size, dtype=np.uint8); nt = ctypes.c_int64(); na = ctypes.c_int64()
    ctx.check(pkg._lib.lib().itjl_abc_accept(ctx.handle, pkg._lib.ptr(dn), dn.size, float(np.nanmedian(dn)), 1000,
                                             pkg._lib.ptr(isacc), ctypes.byref(nt), ctypes.byref(na)))
    assert nt.value == dn.size and na.value == int(np.nansum(dn <= np.nanmedian(dn)))
    assert not isacc[::3].any()


def test_basic_abc_matches_serial_oracle_loop(pkg, ctx):
    """basic_abc (batched, GPU) == the reference's serial loop run on the oracle, given the
    same proposals and the same Philox streams."""
    om, gm = make_plain(pkg, n_t=1200, n_trials=2, seed=3)
    rng = np.random.default_rng(2)
    max_iter, min_acc = 600, 40
    props = rng.uniform(0.05, 2.0, size=(max_iter, 1))
    d_ref_all = cport.abc_distances(om, props, seed=9, generation=5)
    eps = float(np.quantile(d_ref_all, 0.2))
    want = oabc.basic_abc(lambda th, i, g: d_ref_all[i], None, 1, eps, max_iter, min_acc, proposals=props)
    got = pkg.basic_abc(gm, eps, max_iter, min_acc, proposals=props, seed=9, generation=5, ctx=ctx)
    # acceptances can only differ where |d - eps| is inside the fp32 tolerance
    margin = np.abs(d_ref_all - eps) <= dist_tol(d_ref_all)
    assert not margin[: want["n_total"]].any(), "test data too close to the threshold; change the seed"
    assert got["n_total"] == want["n_total"] and got["n_accepted"] == want["n_accepted"]
    assert np.array_equal(got["isaccepted"], want["isaccepted"])
    assert np.array_equal(got["theta_accepted"], want["theta_accepted"])
    assert np.all(np.abs(got["distances"] - want["distances"]) <= dist_tol(want["distances"]))


def test_model_argument_errors(pkg, ctx):
    import ctypes
    d = pkg._lib.ModelDesc()
    h = ctypes.c_void_p()
    tgt = np.ones(10)
    d.kind, d.summary, d.distance, d.update = 0, 0, 0, 0
    d.n_trials, d.n_t, d.dt, d.n_stats, d.target_stats = 2, 1000, 0.002, 10, tgt.ctypes.data
    for field, bad in [("kind", 7), ("summary", 5), ("distance", 3), ("n_trials", 0), ("n_t", 2), ("dt", -1.0),
                       ("n_stats", 0), ("n_stats", 2000)]:
        old = getattr(d, field)
        setattr(d, field, bad)
        rc = pkg._lib.lib().itjl_model_create(ctx.handle, ctypes.byref(d), ctypes.byref(h))
        assert rc == -1, field
        assert len(pkg._lib.lib().itjl_last_error(ctx.handle)) > 0
        setattr(d, field, old)
    d.summary = 1                                              # missing ACF without a mask
    assert pkg._lib.lib().itjl_model_create(ctx.handle, ctypes.byref(d), ctypes.byref(h)) == -1
    d.summary = 0
    assert pkg._lib.lib().itjl_model_create(ctx.handle, ctypes.byref(d), ctypes.byref(h)) == 0
    out = np.empty(4)
    th = np.ones((4, 1))
    assert pkg._lib.lib().itjl_abc_distances(h, pkg._lib.ptr(th), 0, 1, 0, 0, 0, None, None, None,
                                             pkg._lib.ptr(out), None) == -1
    assert pkg._lib.lib().itjl_abc_distances(h, None, 4, 1, 0, 0, 0, None, None, None, pkg._lib.ptr(out), None) == -1
    pkg._lib.lib().itjl_model_destroy(h)


def test_full_size_properties_c5_shape(pkg, ctx):
    """BASELINE config C5 shape (50 trials x 5000 samples, n_lags ~ 480) on a slice of the
    generation: size-independent properties instead of an oracle run."""
    n_t, n_trials = 5000, 50
    data = oou.generate_ou_process(0.3, 1.0, DT, n_t * DT, n_trials, seed=5, n_t=n_t)
    gm = pkg.one_timescale_model(data, _time(n_t), "abc", prior=[pkg.Uniform(0.05, 5)])
    assert 200 < gm.n_lags < 1500
    g = gm.gpu_model(ctx)
    rng = np.random.default_rng(0)
    theta = rng.uniform(0.05, 5.0, size=(592, 1))
    d = g.distances(theta, seed=1, generation=1)
    assert np.isfinite(d).all() and np.all(d >= 0)
    # batching / sharding invariance (what makes results identical across GPU counts)
    parts = [g.distances(theta[lo:hi], seed=1, generation=1, particle_offset=lo)
             for lo, hi in [(0, 74), (74, 300), (300, 592)]]
    assert np.array_equal(np.concatenate(parts), d)
    # the distance is minimised near the generating tau and grows away from it
    grid = np.array([[0.05], [0.15], [0.3], [0.6], [2.5]])
    dg = g.distances(grid, seed=2, generation=1)
    assert np.argmin(dg) == 2 and dg[0] > dg[1] > dg[2] < dg[3] < dg[4]
    # oracle spot check on 2 particles
    om = omodels.one_timescale_model(data, _time(n_t), prior=[omodels.Uniform(0.05, 5)])
    want = cport.abc_distances(om, theta[:2], seed=1, generation=1)
    assert np.all(np.abs(d[:2] - want) <= dist_tol(want))


def test_oscillation_with_missing_model(pkg, ctx):
    """one_timescale_and_osc_with_missing_model, ACF branch: the simulated trial is
    standardised, scaled by data_sd, shifted by data_mean, masked with the data's NaN pattern, and
    comp_ac_time_missing leaves the masked samples at -(mean of the valid samples) - so data_mean
    matters.  Both fused kernels against the oracle, data with a clearly non-zero mean."""
    import os
    dt, n_t, n_trials = 1 / 500, 5000, 6
    t = dt * np.arange(1, n_t + 1)
    data = oou.generate_ou_with_oscillation([0.05, 10.0, 0.9], dt, n_t * dt, n_trials, 1.7, 2.0, seed=4, n_t=n_t)
    rng = np.random.default_rng(3)
    for r in range(n_trials):
        s = rng.integers(0, n_t - 500)
        data[r, s:s + 500] = np.nan
    po = [omodels.Uniform(0.01, 0.5), omodels.Uniform(5, 15), omodels.Uniform(0, 1)]
    pp = [pkg.Uniform(0.01, 0.5), pkg.Uniform(5, 15), pkg.Uniform(0, 1)]
    om = omodels.one_timescale_and_osc_with_missing_model(data, t, prior=po)
    theta = np.array([[0.05, 10.0, 0.9], [0.1, 12.0, 0.5], [0.02, 8.0, 0.2]])
    P = theta.shape[0]
    u0 = rng.standard_normal((P, n_trials)); noise = rng.standard_normal((P, n_trials, n_t))
    phases = rng.uniform(0, 2 * np.pi, size=(P, n_trials))
    want = np.array([omodels.summary_stats(om, omodels.generate_data(om, theta[i], u0=u0[i], noise=noise[i], phases=phases[i]))
                     for i in range(P)])
    for tc in (1, 0):
        gm = pkg.one_timescale_and_osc_with_missing_model(data, t, "abc", prior=pp)
        assert gm.n_lags == om.n_lags and abs(gm.data_mean - om.data_mean) < 1e-12 and abs(gm.data_mean) > 1.0
        assert np.max(np.abs(gm.data_sum_stats - om.data_sum_stats)) < 1e-12
        with ctx.tuned(abc_tc=tc):                    # itjl_tuning: which fused ACF kernel the model gets
            g = gm.gpu_model(ctx)
        assert g.kernel_info()[0] == ("k_abc_acf_tc" if tc == 1 else "k_abc_acf")
        d, stats = g.distances(theta, u0=u0, noise=noise, phases=phases, return_stats=True)
        assert np.max(np.abs(stats - want)) <= ACF_TOL, (tc, np.max(np.abs(stats - want)))
        d_ref = np.mean((want - om.data_sum_stats) ** 2, axis=1)
        assert np.all(np.abs(d - d_ref) <= dist_tol(d_ref))


def test_distance_combined_acf_branch(pkg, ctx):
    """distance_combined = true of one_timescale_model (one_timescale.jl:166-169, 278-320):
    w1 * distance(ACF) + w2 * (data_tau - fit_expdecay(lags, simulated ACF))^2, scored in batches
    (fused kernel -> statistics -> batched exp-decay fit on the device) and through basic_abc's accept
    rule; against the oracle's serial restatement on the same noise."""
    n_t, n_trials = 3000, 4
    data = oou.generate_ou_process(0.3, 1.0, DT, n_t * DT, n_trials, seed=21, n_t=n_t)
    om = omodels.with_combined_distance(
        omodels.one_timescale_model(data, _time(n_t), prior=[omodels.Uniform(0.05, 5)]), weights=(0.3, 0.7))
    gm = pkg.one_timescale_model(data, _time(n_t), "abc", prior=[pkg.Uniform(0.05, 5)], distance_combined=True,
                                 weights=[0.3, 0.7], ctx=ctx)
    assert gm.distance_combined and gm.weights == [0.3, 0.7]
    assert gm.data_tau == pytest.approx(om.data_tau, rel=1e-6)
    rng = np.random.default_rng(8)
    theta = np.array([[0.3], [0.1], [1.0], [2.5], [0.05]])
    P = theta.shape[0]
    u0 = rng.standard_normal((P, n_trials)); noise = rng.standard_normal((P, n_trials, n_t))
    want = np.array([omodels.generate_data_and_reduce(om, theta[i], u0=u0[i], noise=noise[i]) for i in range(P)])
    g = gm.gpu_model(ctx)
    got = g.distances(theta, u0=u0, noise=noise)
    # the tau term amplifies the 1e-5 ACF tolerance by d tau / d ACF ~ tau * n_lags^(1/2)
    assert np.all(np.abs(got - want) <= 2e-3 * np.abs(want) + 1e-6), (got, want)
    # the plain distance is the w1 part
    gp = pkg.one_timescale_model(data, _time(n_t), "abc", prior=[pkg.Uniform(0.05, 5)]).gpu_model(ctx)
    d_plain, stats = gp.distances(theta, u0=u0, noise=noise, return_stats=True)
    tau_sim = pkg.fit_expdecay(gm.lags_freqs, stats, ctx=ctx)
    assert np.allclose(got, 0.3 * d_plain + 0.7 * (gm.data_tau - tau_sim) ** 2, rtol=1e-12, atol=0)
    # batched generation: accept mask / stop index follow the combined distances
    th = rng.uniform(0.05, 3.0, size=(200, 1))
    d_all = g.distances(th, seed=4, generation=1)
    eps = float(np.median(d_all))
    d, isacc, n_total, n_acc = g.generation(th, eps, 30, seed=4, generation=1)
    assert np.array_equal(d, d_all)
    assert n_total == oabc.stop_index(d_all, eps, 30) and n_acc == 30
    assert np.array_equal(isacc[:n_total], (d_all <= eps)[:n_total])


def test_oscillation_with_missing_model_many_long_trials(pkg, ctx):
    """Found by benchmarks/model_fuzz.py: with data_mean != 0 the masked samples of this model sit at one large
    value; with hi*hi tensor-core products only, their common fp16 rounding error put 7e-5 on the ACF at
    50 x 11 000 samples.  The model now always uses the three-term split."""
    dt, n_t, n_trials = 1 / 250, 11000, 50
    t = dt * np.arange(1, n_t + 1)
    data = oou.generate_ou_with_oscillation([0.05, 10.0, 0.9], dt, n_t * dt, n_trials, 0.8, 0.6, seed=49, n_t=n_t)
    rng = np.random.default_rng(3)
    for r in range(n_trials):
        s = int(rng.integers(0, n_t - n_t // 10)); data[r, s:s + n_t // 10] = np.nan
    po = [omodels.Normal(.1, .1), omodels.Normal(10, 5), omodels.Uniform(0, 1)]
    pp = [pkg.Normal(.1, .1), pkg.Normal(10, 5), pkg.Uniform(0, 1)]
    om = omodels.one_timescale_and_osc_with_missing_model(data, t, prior=po)
    gm = pkg.one_timescale_and_osc_with_missing_model(data, t, "abc", prior=pp)
    theta = np.array([[0.05, 10.0, 0.9], [0.1, 12.0, 0.5], [0.02, 8.0, 0.3], [0.2, 11.0, 0.7]])
    want, s_want = cport.abc_distances(om, theta, seed=4, generation=2, return_stats=True)
    got, s_got = gm.gpu_model(ctx).distances(theta, seed=4, generation=2, return_stats=True)
    assert np.max(np.abs(s_got - s_want)) <= ACF_TOL, np.max(np.abs(s_got - s_want))
    assert np.all(np.abs(got - want) <= dist_tol(want))


def test_psd_model_with_more_selected_bins_than_fit_in_shared_memory(pkg, ctx):
    """fs = 250 Hz, n_t > 8192: 0.5 < f < 100 selects 13 042 of the 16 383 bins - their per-particle sums do
    not fit beside the 32 768-point FFT in shared memory (this used to be refused).  They then live in a
    global scratch buffer, 2048 particles per launch; found by benchmarks/model_fuzz.py."""
    fs, n_t, n_trials, P = 250.0, 9501, 2, 2051
    dt = 1 / fs
    t = dt * np.arange(1, n_t + 1)
    data = oou.generate_ou_with_oscillation([0.05, 10.0, 0.9], dt, n_t * dt, n_trials, 0.0, 1.0, seed=3, n_t=n_t)
    po = [omodels.Normal(.1, .1), omodels.Normal(10, 5), omodels.Uniform(0, 1)]
    pp = [pkg.Normal(.1, .1), pkg.Normal(10, 5), pkg.Uniform(0, 1)]
    om = omodels.one_timescale_and_osc_model(data, t, prior=po, summary_method="psd")
    gm = pkg.one_timescale_and_osc_model(data, t, "abc", prior=pp, summary_method="psd")
    assert om.data_sum_stats.size == gm.data_sum_stats.size == 13042
    rng = np.random.default_rng(0)
    theta = np.stack([np.abs(rng.normal(.1, .05, P)) + 0.01, rng.normal(10, 3, P), rng.uniform(0.05, 0.95, P)], axis=1)
    g = gm.gpu_model(ctx)
    got = g.distances(theta, seed=4, generation=2)
    idx = [0, 1, 2047, 2048, 2050]                               # both launches, and their seam
    want = np.array([cport.abc_distances(om, theta[i:i + 1], seed=4, generation=2, particle_offset=i)[0] for i in idx])
    assert np.all(np.abs(got[idx] - want) <= 1e-5 * want)
    got5, stats5 = g.distances(theta[:5], seed=4, generation=2, return_stats=True)
    _, s_or = cport.abc_distances(om, theta[:5], seed=4, generation=2, return_stats=True)
    assert np.array_equal(got5, got[:5])
    assert np.max(np.abs(stats5 - s_or) / s_or.max(axis=1, keepdims=True)) <= 1e-5
