"""GPU parity of the PSD-side reducers: batched Lorentzian knee / oscillation peak / FOOOF loop
(itjl_fit_knee), acw(...; acwtypes = :knee) on complete data and on data with gaps (itjl_acw_knee), and the combined
PSD distances of one_timescale_model and one_timescale_and_osc_model - against oracle/knee.py, whose minimiser is
SciPy's Nelder-Mead + an exact weighted-median polish, not the device's search."""
import numpy as np
import pytest

from oracle import knee as oknee
from oracle import models as omodels
from oracle import ou as oou
from oracle import summary as osum

pytestmark = pytest.mark.gpu


def _spectra(rng, n_rows, f, noise):
    rows = []
    for r in range(n_rows):
        amp, k = 10 ** rng.uniform(-1, 2), rng.uniform(2.0, 40.0)
        p = amp / (1 + (f / k) ** 2)
        if r % 2:
            p = p + 0.3 * amp * np.exp(-(f - rng.uniform(15, 60)) ** 2 / (2 * rng.uniform(1.0, 3.0) ** 2))
        rows.append(p * (1.0 + noise * rng.standard_normal(f.size)))
    return np.array(rows)


def _objective(p, f, u):
    return float(np.mean(np.abs(oknee.lorentzian(f, u) - p)))


@pytest.mark.parametrize("noise", [0.0, 0.05, 0.3])
def test_find_knee_frequency_batch(pkg, ctx, noise):
    rng = np.random.default_rng(int(noise * 100))
    f = np.arange(1, 481) * 0.25
    P = _spectra(rng, 12, f, noise)
    got, peaks = pkg.utils.find_knee_frequency(P, f, ctx=ctx, return_peak=True)
    for r in range(P.shape[0]):
        want = oknee.find_knee_frequency(P[r], f)
        # the device's answer is at least as good a minimiser of the reference's objective as the oracle's ...
        assert _objective(P[r], f, got[r]) <= _objective(P[r], f, want) * (1 + 1e-9) + 1e-12 * P[r].max()
        # ... and it is the same point (an L1 objective is kinked: compared to 1e-6 relative, 1e-4 with heavy noise)
        tol = 1e-6 if noise < 0.2 else 1e-4
        assert abs(got[r, 1] - want[1]) <= tol * want[1] and abs(got[r, 0] - want[0]) <= tol * want[0], (r, got[r], want)
        if noise > 0 or r % 2:              # (the residual of an exact fit to a pure Lorentzian is rounding noise: no peak to compare)
            wp = oknee.find_oscillation_peak(P[r] - oknee.lorentzian(f, want), f, f[0], f[-1])
            assert (np.isnan(wp) and np.isnan(peaks[r])) or peaks[r] == wp
    one = pkg.utils.find_knee_frequency(P[3], f, ctx=ctx)
    assert one[1] == got[3, 1]
    g = pkg.utils.lorentzian_initial_guess(P, f, ctx=ctx)
    for r in range(P.shape[0]):
        assert np.allclose(g[r], oknee.lorentzian_initial_guess(P[r], f), rtol=1e-14)


def test_reference_edge_cases_and_peak_finder(pkg, ctx):
    freqs = np.arange(0, 1001) * 0.1 / 1000.0
    assert np.isclose(pkg.utils.find_knee_frequency(np.ones(freqs.size), freqs, ctx=ctx)[1], freqs[-1])      # test_utils.jl:76-77
    psd = 100.0 / (1 + (freqs[1:] / 0.015) ** 2)
    assert abs(pkg.utils.find_knee_frequency(psd, freqs[1:], ctx=ctx)[1] - 0.015) < 1e-3 * 0.015            # :11-19
    s = np.array([5.0, 1.0, 4.0, 3.5, 6.0, 0.0, 2.0, 1.5])
    assert pkg.utils.find_oscillation_peak(s, np.arange(8.0), 0.0, 7.0, ctx=ctx) == 4.0
    assert np.isnan(pkg.utils.find_oscillation_peak(np.arange(8.0)[::-1].copy(), np.arange(8.0), 0.0, 7.0, ctx=ctx))
    rng = np.random.default_rng(1)
    f = np.arange(1, 2001) * 0.05
    R = rng.standard_normal((9, f.size)).cumsum(axis=1)
    got = pkg.utils.find_oscillation_peak(R, f, 3.0, 77.0, ctx=ctx)
    for r in range(R.shape[0]):
        w = oknee.find_oscillation_peak(R[r], f, 3.0, 77.0)
        assert (np.isnan(w) and np.isnan(got[r])) or got[r] == w


def test_fooof_fit_batch(pkg, ctx):
    rng = np.random.default_rng(7)
    f = np.arange(1, 601) * 0.2
    P = _spectra(rng, 8, f, 0.02)
    for max_peaks in (1, 3):
        got = pkg.utils.fooof_fit(P, f, min_freq=1.0, max_freq=100.0, oscillation_peak=True, max_peaks=max_peaks,
                                  return_only_knee=True, ctx=ctx)
        for r in range(P.shape[0]):
            want = oknee.fooof_fit(P[r], f, 1.0, 100.0, True, max_peaks, True)
            assert abs(got[r] - want) <= 2e-4 * want, (max_peaks, r, got[r], want)
    got = pkg.utils.fooof_fit(P, f, min_freq=1.0, max_freq=100.0, oscillation_peak=False, ctx=ctx)
    for r in range(P.shape[0]):
        assert abs(got[r] - oknee.fooof_fit(P[r], f, 1.0, 100.0, False)) <= 1e-6 * got[r]


def _oracle_acw_knee(data, fs, freqlims, oscillation_peak, max_peaks, average):
    """acw.jl:242-270 on (trials, time)."""
    dt = 1.0 / fs
    mask = np.isnan(data)
    if mask.any():
        t = dt * np.arange(1, data.shape[1] + 1)
        psd, freqs = osum.comp_psd_lombscargle(t, data, mask, dt)
    else:
        psd, freqs = osum.comp_psd_periodogram(data, fs)
    if average:
        psd = psd.mean(axis=0, keepdims=True)
    lo, hi = (freqs[0], freqs[-1]) if freqlims is None else freqlims
    k = np.array([oknee.fooof_fit(p, freqs, lo, hi, oscillation_peak, max_peaks, True) for p in psd])
    return oknee.tau_from_knee(k), psd, freqs


@pytest.mark.parametrize("gaps,average,osc_peak", [(False, False, True), (False, True, False), (True, False, False), (True, True, True)])
def test_acw_knee(pkg, ctx, gaps, average, osc_peak):
    """test_acw.jl:73 / :261-280 shapes: OU trials, fs = 100, freqlims (0.1, 5) with gaps."""
    fs, n_t, n_trials, tau = 100.0, 2000, 6, 0.3
    data = oou.generate_ou_process(tau, 1.0, 1 / fs, n_t / fs, n_trials, seed=12, n_t=n_t)
    freqlims = None
    if gaps:
        rng = np.random.default_rng(3)
        for r in range(n_trials):
            s = rng.integers(0, n_t - 200)
            data[r, s:s + 150] = np.nan
        freqlims = (0.1, 5.0) if osc_peak else None           # default freqlims = the whole Lomb-Scargle axis, Nyquist included
    want_tau, want_psd, want_f = _oracle_acw_knee(data, fs, freqlims, osc_peak, 1, average)
    got = pkg.acw(data, fs, acwtypes="knee", freqlims=freqlims, average_over_trials=average, oscillation_peak=osc_peak,
                  max_peaks=1, ctx=ctx)
    assert np.allclose(got.freqs, want_f, rtol=1e-12)
    assert np.all(np.isfinite(got.psd))
    assert np.max(np.abs(np.atleast_2d(got.psd) - want_psd) / want_psd.max()) < 1e-9
    assert np.allclose(np.atleast_1d(got.acw_results), want_tau, rtol=5e-4), (got.acw_results, want_tau)
    # the knee recovers the timescale (test_acw.jl: within a factor of two for six 20 s trials)
    assert np.all(np.abs(np.log(np.atleast_1d(got.acw_results) / tau)) < np.log(3.0))
    both = pkg.acw(data, fs, acwtypes=["acw50", "knee"], freqlims=freqlims, average_over_trials=average,
                   oscillation_peak=osc_peak, ctx=ctx)
    assert both.acwtypes == ["acw50", "knee"] and np.allclose(both.acw_results[1], got.acw_results)


def test_combined_psd_distances(pkg, ctx):
    """distance_combined = true on the PSD summaries: one_timescale_model (weights [w1, w2], knee timescale) and
    one_timescale_and_osc_model (three weights, knee timescale + oscillation peak), informed priors included."""
    dt, n_t, n_trials = 1 / 500, 5000, 6
    t = dt * np.arange(1, n_t + 1)
    data = oou.generate_ou_process(0.05, 1.0, dt, n_t * dt, n_trials, seed=3, n_t=n_t)
    om = omodels.with_combined_distance(omodels.one_timescale_model(data, t, summary_method="psd"), (0.7, 0.3))
    gm = pkg.one_timescale_model(data, t, "abc", summary_method="psd", distance_combined=True, weights=[0.7, 0.3], ctx=ctx)
    assert abs(gm.prior[0].mu - om.prior[0].mu) <= 1e-5 * om.prior[0].mu and gm.prior[0].sigma == 20.0
    assert abs(gm.data_tau - om.data_tau) <= 1e-5 * om.data_tau
    theta = np.array([[0.05], [0.02], [0.2]])
    got = gm.gpu_model(ctx).distances(theta, seed=2, generation=1)
    for i in range(3):
        x = omodels.generate_data(om, theta[i], seed=2, generation=1, particle=i)
        want = omodels.distance_function(om, omodels.summary_stats(om, x), om.data_sum_stats)
        assert abs(got[i] - want) <= 1e-3 * want, (i, got[i], want)

    data = oou.generate_ou_with_oscillation([0.05, 10.0, 0.7], dt, n_t * dt, n_trials, 0.0, 1.0, seed=11, n_t=n_t)
    om = omodels.with_combined_distance(omodels.one_timescale_and_osc_model(data, t, summary_method="psd"), (0.5, 0.3, 0.2))
    gm = pkg.one_timescale_and_osc_model(data, t, "abc", summary_method="psd", distance_combined=True, weights=[0.5, 0.3, 0.2],
                                         ctx=ctx)
    for a, b in zip(gm.prior, om.prior):
        assert type(a).__name__ == type(b).__name__
    assert abs(gm.prior[0].mu - om.prior[0].mu) <= 1e-9 * abs(om.prior[0].mu) and gm.prior[1].mu == om.prior[1].mu
    assert abs(gm.data_tau - om.data_tau) <= 1e-5 * om.data_tau and gm.data_osc == om.data_osc
    with pytest.raises(ValueError):
        pkg.one_timescale_and_osc_model(data, t, "abc", summary_method="psd", distance_combined=True, ctx=ctx)    # two weights
    theta = np.array([[0.05, 10.0, 0.7], [0.1, 14.0, 0.4], [0.02, 8.0, 0.9]])
    got = gm.gpu_model(ctx).distances(theta, seed=5, generation=3)
    for i in range(3):
        x = omodels.generate_data(om, theta[i], seed=5, generation=3, particle=i)
        want = omodels.distance_function(om, omodels.summary_stats(om, x), om.data_sum_stats)
        assert (np.isnan(want) and np.isnan(got[i])) or abs(got[i] - want) <= 1e-3 * want, (i, got[i], want)


def test_combined_distance_on_the_device_equals_the_host_combination(pkg, ctx):
    """itjl_abc_distances_combined (statistics stay on the device) against the step-by-step path (statistics to the
    host, itjl_fit_knee / itjl_fit_expdecay, NumPy combination) for the three combined distances."""
    dt, n_t, n_trials = 1 / 500, 5000, 4
    t = dt * np.arange(1, n_t + 1)
    rng = np.random.default_rng(3)
    d1 = oou.generate_ou_process(0.05, 1.0, dt, n_t * dt, n_trials, seed=3, n_t=n_t)
    d2 = oou.generate_ou_with_oscillation([0.05, 10.0, 0.7], dt, n_t * dt, n_trials, 0.0, 1.0, seed=11, n_t=n_t)
    models = [
        (pkg.one_timescale_model(d1, t, "abc", summary_method="psd", distance_combined=True, weights=[0.7, 0.3], ctx=ctx),
         rng.uniform(0.01, 0.5, size=(40, 1))),
        (pkg.one_timescale_model(d1, t, "abc", distance_combined=True, weights=[0.4, 0.6], ctx=ctx), rng.uniform(0.01, 0.5, size=(40, 1))),
        (pkg.one_timescale_and_osc_model(d2, t, "abc", summary_method="psd", distance_combined=True, weights=[0.5, 0.3, 0.2], ctx=ctx),
         np.stack([rng.uniform(0.02, 0.2, 40), rng.uniform(5, 20, 40), rng.uniform(0.1, 0.95, 40)], axis=1)),
    ]
    for gm, theta in models:
        g = gm.gpu_model(ctx)
        fused = g.distances(theta, seed=7, generation=2)
        plain, stats = g.distances(theta, seed=7, generation=2, return_stats=True)     # host combination of the same statistics
        assert np.allclose(fused, plain, rtol=1e-12, atol=0.0, equal_nan=True), np.nanmax(np.abs(fused - plain) / plain)
        d, isacc, n_total, n_acc = g.generation(theta, float(np.nanmedian(fused)), 10, seed=7, generation=2)
        assert np.array_equal(d[:n_total], fused[:n_total], equal_nan=True) and n_acc == 10
