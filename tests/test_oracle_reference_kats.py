"""Pins the CPU oracle against every known-answer / property test the reference holds for
the hot path.  The reference ships no golden vectors and cannot be
executed here (no Julia), so these KATs - restated with the reference test's own tolerance -
plus tests/golden/ are what the oracle is anchored on."""
import math

import numpy as np
import pytest

from oracle import abc, acw, acwred, distances, models, ou, philox, summary


def test_philox_random123_kat():
    # Random123 kat_vectors, philox4x32-10
    kats = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
            ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
            ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
             (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
    for ctr, key, want in kats:
        got = philox.philox4x32_10(*ctr, *key)
        assert tuple(int(g) for g in got) == want
    # the stream of this project uses 7 rounds of the same round function (kat_vectors: philox4x32 7, all-zero input)
    assert philox.ROUNDS == 7
    got = philox.philox4x32(0, 0, 0, 0, 0, 0)
    assert tuple(int(g) for g in got) == (0x5f6fb709, 0x0d893f64, 0x4f121f81, 0x4f730a48)


def test_philox_normals_moments():
    z = philox.noise_normals(7, 3, 11, 2, 400000)
    assert abs(z.mean()) < 5e-3 and abs(z.std() - 1) < 5e-3
    assert abs((z ** 4).mean() - 3.0) < 0.05
    assert abs(np.corrcoef(z[:-1], z[1:])[0, 1]) < 5e-3


# ---- test/test_distances.jl:4-32 ---------------------------------------------------
def test_linear_distance_kat():
    x, y = [1.0, 2.0, 3.0], [1.1, 2.1, 3.1]
    d = distances.linear_distance(x, y)
    assert d >= 0 and distances.linear_distance(x, x) == 0
    assert d == pytest.approx(distances.linear_distance(y, x))
    assert d == pytest.approx(0.01, abs=1e-10)


def test_logarithmic_distance_kat():
    x, y = np.array([1.0, 2.0, 3.0]), np.array([1.1, 2.2, 3.3])
    d = distances.logarithmic_distance(x, y)
    assert d >= 0 and distances.logarithmic_distance(x, x) == pytest.approx(0)
    assert d == pytest.approx(np.mean((np.log(x) - np.log(y)) ** 2))


# ---- test/test_summary_stats.jl:20-36, 57-101 ----------------------------------------
def test_comp_ac_fft_sine():
    t = np.arange(0, 10.0001, 0.1)
    ac = summary.comp_ac_fft_vec(np.sin(t), n_lags=100)
    assert len(ac) == len(t) - 1
    assert ac[0] == pytest.approx(1.0, abs=0.1) and np.all(ac <= 1.0 + 1e-12)
    period = round(2 * math.pi / 0.1)
    assert np.diff(ac)[period - 1] == pytest.approx(0.0, abs=0.1)


def test_three_acf_methods_agree_on_complete_data():
    dt, T, tau, n_trials = 0.01, 10.0, 1.0, 20
    x = ou.generate_ou_process(tau, 1.0, dt, T, n_trials, seed=123)
    max_lags = 100
    ac_fft = summary.comp_ac_fft(x, max_lags).mean(axis=0)
    ac_miss = summary.comp_ac_time_missing(x, max_lags).mean(axis=0)
    ac_dir = np.mean([summary.comp_ac_direct_vec(r, max_lags) for r in x], axis=0)
    assert np.max(np.abs(ac_dir - ac_miss)) < 1e-13      # reference: < 1e-3 (observed 2e-16)
    assert np.max(np.abs(ac_miss - ac_fft)) < 1e-13
    lags = np.arange(max_lags) * dt
    assert np.corrcoef(ac_miss, np.exp(-lags / tau))[0, 1] > 0.95


def test_acf_with_missing_values():
    dt, T, tau = 0.01, 10.0, 1.0
    x = ou.generate_ou_process(tau, 1.0, dt, T, 20, seed=5)
    ac_full = summary.comp_ac_time_missing(x, 100).mean(axis=0)
    rng = np.random.default_rng(0)
    xm = x.copy()
    xm[0, rng.integers(0, x.shape[1], x.shape[1] // 5)] = np.nan
    ac = summary.comp_ac_time_missing(xm, 100).mean(axis=0)
    assert ac[0] == pytest.approx(1.0, abs=0.1) and ac[-1] < ac[0]
    assert np.all(np.diff(ac[:50]) < 0)
    xs = x.copy()
    xs[0, rng.integers(0, x.shape[1])] = np.nan
    assert np.max(np.abs(summary.comp_ac_time_missing(xs, 100).mean(axis=0) - ac_full)) < 1e-3


def test_missing_acf_masked_samples_become_minus_mean():
    # summary.jl:466-471: NaN -> 0, then the valid-sample mean is subtracted from EVERY entry
    x = np.array([1.0, np.nan, 3.0, 5.0, np.nan, 2.0])
    v = np.array([1.0, 0.0, 3.0, 5.0, 0.0, 2.0])
    z = v - v.sum() / 4
    want = np.array([np.dot(z[:6 - k], z[k:]) for k in range(4)]) / np.dot(z, z)
    assert np.allclose(summary.comp_ac_time_missing_vec(x, 4), want, rtol=0, atol=1e-15)


def test_psd_adfriendly_peaks_and_scaling():
    fs = 100.0
    t = np.arange(0, 1.0001, 1 / fs)
    rng = np.random.default_rng(1)
    x = np.sin(2 * np.pi * 10 * t) + 0.5 * np.sin(2 * np.pi * 25 * t) + 0.1 * rng.standard_normal(t.size)
    p, f = summary.comp_psd_adfriendly_vec(x, fs)
    n2 = 2 * summary._ac_next_pow_two(t.size)
    assert p.shape == (n2 // 2 - 1,) and f[0] == pytest.approx(fs / n2)
    peaks = [i for i in range(1, len(p) - 1) if p[i] > p[i - 1] and p[i] > p[i + 1]]
    top = sorted(peaks, key=lambda i: -p[i])[:2]
    assert sorted(round(f[i]) for i in top) == [10, 25]
    # Parseval for the windowed, padded signal: sum |X|^2 = n2 * sum (x w)^2
    w = summary.hamming_window(t.size)
    xw = (x - x.mean()) * w
    full = np.abs(np.fft.fft(xw, n2)) ** 2
    assert full[1:n2 // 2] / (fs * np.sum(w ** 2)) == pytest.approx(p)


# ---- test/test_ou_process.jl:18-44 ----------------------------------------------------
def test_ou_process_statistics():
    tau, D, dt, T, n_trials = 0.5, 2.0, 0.001, 10.0, 10
    x = ou.generate_ou_process(tau, D, dt, T, n_trials, seed=123)
    assert x.shape == (n_trials, int(T / dt))
    assert abs(x.mean(axis=1).mean()) < 0.1
    assert np.mean(np.abs(x.std(axis=1, ddof=1) - D)) < 0.1
    ac = np.mean([np.corrcoef(r[:-1], r[1:])[0, 1] for r in x])
    assert abs(ac - math.exp(-dt / tau)) < 0.01


def test_ou_exact_update_passes_where_em_fails():
    # lag-1 autocorrelation = exp(-dt/tau) also when dt/tau is not small
    tau, dt, T = 1.0, 1.0, 20000.0
    x = ou.generate_ou_process(tau, 1.0, dt, T, 4, seed=1)
    ac = np.mean([np.corrcoef(r[:-1], r[1:])[0, 1] for r in x])
    assert abs(ac - math.exp(-1.0)) < 0.01
    xe = ou.generate_ou_process(tau, 1.0, dt, T, 4, seed=1, update="em")
    ace = np.mean([np.corrcoef(r[:-1], r[1:])[0, 1] for r in xe])
    assert abs(ace - math.exp(-1.0)) > 0.1                    # EM: AR coefficient 1 - dt/tau = 0


def test_ou_edge_cases_and_osc():
    assert np.isfinite(ou.generate_ou_process(0.1, 10.0, 0.01, 1.0, 10, seed=1)).all()
    assert np.isfinite(ou.generate_ou_process(1000.0, 0.001, 1.0, 100.0, 10, seed=1)).all()
    assert ou.generate_ou_process(20.0, 0.05, 1.0, 100.0, 1, seed=1).shape[0] == 1
    y = ou.generate_ou_with_oscillation([0.05, 10.0, 0.95], 1 / 500, 4.0, 5, 2.0, 3.0, seed=4)
    assert y.shape == (5, 2000)
    assert np.allclose(y.mean(axis=1), 2.0) and np.allclose(y.std(axis=1, ddof=1), 3.0)
    assert ou.clamp_coeff(-0.5) == 1e-4 and ou.clamp_coeff(1.5) == 1 - 1e-4 and ou.clamp_coeff(0.3) == 0.3


# ---- test/test_utils.jl:119-197, 335-349 -------------------------------------------------
@pytest.mark.parametrize("tau", [0.1, 0.5, 2.0, 5.0, 100.0])
def test_fit_expdecay_recovers_tau(tau):
    lags = np.linspace(0.0, 10.0 * tau if tau < 50 else 100.0, 200)
    assert acwred.fit_expdecay(lags, np.exp(-lags / tau)) == pytest.approx(tau, rel=1e-6)


def test_fit_expdecay_noisy_and_is_argmin():
    from scipy.optimize import minimize_scalar
    rng = np.random.default_rng(3)
    lags = np.arange(0.0, 50.0)
    acf = np.exp(-lags / 7.0) + 0.05 * rng.standard_normal(lags.size)
    tau = acwred.fit_expdecay(lags, acf)
    assert tau == pytest.approx(7.0, rel=0.1)
    ref = minimize_scalar(lambda u: acwred.expdecay_residual(math.exp(u), lags, acf),
                          bounds=(math.log(0.5), math.log(100.0)), method="bounded",
                          options=dict(xatol=1e-12))
    assert tau == pytest.approx(math.exp(ref.x), rel=1e-5)


def test_acw50_on_exponential():
    lags = np.arange(0.0, 50.0)
    acf = np.exp(-lags / 5.0)
    a50 = acwred.acw50(lags, acf)
    assert a50 == 4.0                                         # exp(-3/5)=0.549 > .5 >= exp(-4/5)
    assert acwred.tau_from_acw50(a50) == pytest.approx(-4.0 / math.log(0.5))
    assert math.isnan(acwred.acw50(lags, np.ones_like(lags)))
    assert acwred.acw0(lags, acf - 0.3) == 7.0
    assert acwred.acweuler(lags, acf) == 5.0


def test_romberg_kats():
    assert acwred.acw_romberg(1.0, [1.0, 0.5, 0.0]) == pytest.approx(1.0, rel=0.01)
    assert acwred.acw_romberg(0.5, np.ones(5)) == pytest.approx(2.0, rel=0.01)
    x = np.linspace(0, 1, 65)
    assert acwred.acw_romberg(x[1] - x[0], np.exp(x)) == pytest.approx(math.e - 1, rel=1e-12)
    x = np.linspace(0, 1, 13)                                # non power-of-two length
    assert acwred.acw_romberg(x[1] - x[0], np.exp(x)) == pytest.approx(math.e - 1, rel=1e-9)
    with pytest.raises(ValueError):
        acwred.romberg(1.0, [1.0])


# ---- test/test_acw.jl:128-143, 285-345 ----------------------------------------------------
def test_acw_front_end_properties():
    fs, tau_s, tau_l = 100.0, 0.3, 1.0
    short = ou.generate_ou_process(tau_s, 1.0, 1 / fs, 50.0, 10, seed=2)
    long_ = ou.generate_ou_process(tau_l, 1.0, 1 / fs, 50.0, 10, seed=2, particle=1)
    rs, rl = acw.acw(short, fs), acw.acw(long_, fs)
    assert np.mean(rl["acw50"]) > np.mean(rs["acw50"])
    assert np.mean(rs["acw50"]) == pytest.approx(tau_s * math.log(2), rel=0.5)
    assert np.mean(rs["tau"]) == pytest.approx(tau_s, rel=0.5)
    assert np.all(rs["auc"] > 0) and np.all(rs["auc"] < 50.0)
    assert rs["acf"].shape == (10, rs["n_lags"]) and rs["lags"].shape == (rs["n_lags"],)
    assert rs["n_lags"] == math.ceil(1.1 * rs["k0"].max())
    nan = short.copy()
    nan[0, 100:140] = np.nan
    rn = acw.acw(nan, fs)
    assert np.isfinite(rn["acw50"]).all()
    avg = acw.acw(short, fs, average_over_trials=True)
    assert avg["acf"].shape[0] == 1


# ---- test/test_abc.jl:11-48, 66-89 ---------------------------------------------------------
def test_abc_bookkeeping():
    res = abc.basic_abc(lambda th, i, g: 0.0, lambda: np.array([1.0, 2.0]), 2, epsilon=1.0,
                        max_iter=100, min_accepted=10)
    assert res["n_total"] == 10 and res["n_accepted"] == 10
    assert res["theta_accepted"].shape == (10, 2) and res["samples"].shape == (100, 2)
    res = abc.basic_abc(lambda th, i, g: float("nan"), lambda: np.array([1.0]), 1, epsilon=1.0,
                        max_iter=50, min_accepted=10)
    assert res["n_total"] == 50 and res["n_accepted"] == 0
    assert abc.effective_sample_size(np.ones(3)) == pytest.approx(3.0)
    x = np.array([[1.0, 2.0], [2.0, 3.0], [3.0, 4.0]])
    wc = abc.weighted_covar(x, np.ones(3))
    assert wc.shape == (2, 2) and np.allclose(wc, np.cov(x, rowvar=False))
    prior = [models.Uniform(0, 10), models.Uniform(0, 10)]
    w = abc.calc_weights(x, x + 0.1, np.eye(2), np.ones(3) / 3, prior)
    assert w.sum() == pytest.approx(1.0) and np.all(w > 0)
    w1 = abc.calc_weights(x[:, :1], x[:, :1] + 0.1, np.eye(1), np.ones(3) / 3, prior[:1])
    assert w1.sum() == pytest.approx(1.0)


def test_select_epsilon_and_alpha():
    d = np.concatenate([np.linspace(0, 1, 101), [np.nan, 50.0]])
    assert abc.select_epsilon(d, 1.0) == pytest.approx(0.5)
    assert abc.percentile_type7([1, 2, 3, 4], 25) == pytest.approx(1.75)
    a = abc.compute_adaptive_alpha(5, 0.5, 0.01, total_iterations=30)
    assert a == pytest.approx(min(0.9, (0.9 * (1 - 5 / 30) + 0.1 * 5 / 30) * 1.5))
    e = abc.select_epsilon(d, 0.8, current_acc_rate=0.5, iteration=5, total_iterations=30)
    assert e == pytest.approx(max(0.25, 0.8 * (1 - a)))
    assert abc.select_epsilon(np.array([np.nan]), 0.3) == 0.3


# ---- test/test_one_timescale.jl:75-127 (model protocol) ------------------------------------
def test_model_protocol_shapes():
    dt, n = 0.002, 2500
    t = dt * np.arange(1, n + 1)
    data = ou.generate_ou_process(0.3, 1.0, dt, t[-1], 3, seed=9, n_t=n)
    m = models.one_timescale_model(data, t, prior=[models.Uniform(0.05, 2)])
    assert m.data_sum_stats.shape == (m.n_lags,) and m.data_sum_stats[0] == pytest.approx(1.0)
    synth = models.generate_data(m, [0.3], seed=1, generation=0, particle=0)
    assert synth.shape == data.shape
    assert abs(synth.mean()) < 0.1 and abs(synth.std() - m.data_sd) < 0.1
    s = models.summary_stats(m, synth)
    assert s.shape == (m.n_lags,) and s[0] == pytest.approx(1.0) and np.all(np.abs(s) <= 1 + 1e-12)
    assert models.distance_function(m, s, s) == pytest.approx(0.0, abs=1e-10)
    assert models.distance_function(m, s, m.data_sum_stats) > 0
    assert isinstance(m.prior[0], models.Uniform)
    mi = models.one_timescale_model(data, t)                  # informed prior Normal(tau_hat, 20)
    assert isinstance(mi.prior[0], models.Normal) and mi.prior[0].sigma == 20.0
    assert mi.prior[0].mu == pytest.approx(0.3, rel=0.7)


def test_osc_with_missing_model_acf_branch():
    """test/test_one_timescale_and_osc_with_missing.jl:74-100 (generate_data keeps the NaN pattern) and
    :209-268 (ACF branch: distance to self is 0, to a faster oscillation > 0), plus the masked-sample
    quirk of comp_ac_time_missing with the model's data_mean (summary.jl:466-471)."""
    from oracle import models as omodels, ou as oou, summary as osum
    dt, T, n_trials = 1.0, 1000.0, 5
    time = dt * np.arange(1, 1001)
    d1 = oou.generate_ou_with_oscillation([20.0, 0.01, 0.5], dt, T, n_trials, 0.0, 1.0, seed=1)
    d2 = oou.generate_ou_with_oscillation([20.0, 0.05, 0.5], dt, T, n_trials, 0.0, 1.0, seed=1)
    mask = np.random.default_rng(0).random(d1.shape) < 0.1
    d1[mask] = np.nan; d2[mask] = np.nan
    prior = [omodels.Uniform(1, 100), omodels.Uniform(0.001, 0.1), omodels.Uniform(0, 1)]
    m = omodels.one_timescale_and_osc_with_missing_model(d1, time, summary_method="acf", n_lags=100,
                                                         distance_method="linear", prior=prior)
    assert m.n_lags == 100 and m.missing_mask.shape == d1.shape and np.array_equal(m.missing_mask, mask)
    s1, s2 = omodels.summary_stats(m, d1), omodels.summary_stats(m, d2)
    assert np.allclose(s1, m.data_sum_stats, atol=0, rtol=0)
    assert omodels.distance_function(m, s1, s1) == pytest.approx(0.0, abs=1e-10)
    assert omodels.distance_function(m, s1, s2) > 0.0
    sim = omodels.generate_data(m, np.array([20.0, 0.02, 0.5]), seed=3, generation=0, particle=0)
    assert sim.shape == d1.shape and np.all(np.isnan(sim[mask])) and not np.any(np.isnan(sim[~mask]))
    # masked samples of a series with mean mu enter the lag sums as -(mean of the valid samples)
    x = np.array([[3.0, np.nan, 5.0, 4.0, np.nan, 6.0]])
    z = np.array([3.0, 0.0, 5.0, 4.0, 0.0, 6.0]) - 4.5
    want = np.array([np.dot(z[:6 - k], z[k:]) for k in range(3)]) / np.dot(z, z)
    assert np.allclose(osum.comp_ac_time_missing(x, 3)[0], want, atol=1e-15)
