"""Seeded random shapes through both fused ACF kernels against the fp64 C oracle.

The shapes are drawn the way benchmarks/tc_fuzz.py and benchmarks/model_fuzz.py draw them; those runs found
the accumulator stagnation of the CUDA-core kernel, the shared rounding residual of masked samples in the
tensor-core kernel and the lag-window / tail seams - this keeps a fixed sample of them in the suite."""
import os

import numpy as np
import pytest

from oracle import cport
from oracle import models as omodels
from oracle import ou as oou

pytestmark = pytest.mark.gpu

ACF_TOL = 1e-5


def _cases():
    rng = np.random.default_rng(2024)
    out = []
    for case in range(14):
        n_t = int(rng.integers(300, 12000))
        n_trials = int(rng.choice([1, 2, 3, 5, 7, 13, 30]))
        n_lags = int(rng.integers(2, min(n_t, 1537) + 1))
        missing = bool(rng.integers(0, 3) == 0)
        out.append((case, n_t, n_trials, n_lags, missing))
    return out


@pytest.mark.parametrize("case,n_t,n_trials,n_lags,missing", _cases())
def test_random_shape_both_kernels_against_fp64_oracle(pkg, ctx, case, n_t, n_trials, n_lags, missing):
    dt = 0.002
    data = oou.generate_ou_process(0.3, 1.0, dt, n_t * dt, n_trials, seed=case + 1, n_t=n_t)
    if missing:
        rng = np.random.default_rng(case)
        for r in range(n_trials):
            s = int(rng.integers(0, n_t - n_t // 10))
            data[r, s:s + n_t // 10] = np.nan
    t = dt * np.arange(1, n_t + 1)
    if missing:
        om = omodels.one_timescale_with_missing_model(data, t, prior=[omodels.Uniform(0.05, 5)], n_lags=n_lags)
        mk = lambda: pkg.one_timescale_with_missing_model(data, t, "abc", prior=[pkg.Uniform(0.05, 5)], n_lags=n_lags)
    else:
        om = omodels.one_timescale_model(data, t, prior=[omodels.Uniform(0.05, 5)], n_lags=n_lags)
        mk = lambda: pkg.one_timescale_model(data, t, "abc", prior=[pkg.Uniform(0.05, 5)], n_lags=n_lags)
    theta = np.random.default_rng(100 + case).uniform(0.05, 5.0, size=(6, 1))
    d_or, s_or = cport.abc_distances(om, theta, seed=3, generation=1, return_stats=True)
    for tc in (1, 0):
        with ctx.tuned(abc_tc=tc):
            g = mk().gpu_model(ctx)
        d, s = g.distances(theta, seed=3, generation=1, return_stats=True)
        err = np.max(np.abs(s - s_or))
        assert err <= ACF_TOL, (g.kernel_info()[0], n_t, n_trials, n_lags, missing, err)
        assert np.all(np.abs(d - d_or) <= 2 * np.sqrt(d_or) * ACF_TOL + ACF_TOL ** 2 + 1e-5 * d_or)


def _model_cases():
    rng = np.random.default_rng(4242)
    out = []
    for case in range(10):
        out.append((case, ["osc_acf", "osc_psd", "osc_acf_missing", "plain_em"][int(rng.integers(0, 4))], int(rng.integers(400, 12000)),
                    int(rng.choice([1, 2, 5, 20, 50])), float(rng.choice([250.0, 500.0, 1000.0])), bool(rng.integers(0, 2))))
    return out


@pytest.mark.parametrize("case,kind,n_t,n_trials,fs,shifted", _model_cases())
def test_random_shape_oscillation_and_em_models_against_fp64_oracle(pkg, ctx, case, kind, n_t, n_trials, fs, shifted):
    """A fixed sample of benchmarks/model_fuzz.py: oscillation model with the ACF and the PSD summary, with
    missing data (data_mean != 0 moves the masked samples), and the Euler-Maruyama update."""
    dt = 1.0 / fs
    t = dt * np.arange(1, n_t + 1)
    rng = np.random.default_rng(500 + case)
    po = [omodels.Normal(.1, .1), omodels.Normal(10, 5), omodels.Uniform(0, 1)]
    pp = [pkg.Normal(.1, .1), pkg.Normal(10, 5), pkg.Uniform(0, 1)]
    rel = False
    if kind == "plain_em":
        data = oou.generate_ou_process(0.3, 1.0, dt, n_t * dt, n_trials, seed=case + 1, n_t=n_t)
        om = omodels.one_timescale_model(data, t, prior=[omodels.Uniform(0.05, 5)], update="em")
        gm = pkg.one_timescale_model(data, t, "abc", prior=[pkg.Uniform(0.05, 5)], update="em")
        theta = rng.uniform(0.05, 5.0, size=(5, 1)); tol = 1e-5
    else:
        dm, ds = (float(rng.normal(0, 1)), float(rng.uniform(0.5, 3))) if shifted else (0.0, 1.0)
        data = oou.generate_ou_with_oscillation([0.05, 10.0, 0.9], dt, n_t * dt, n_trials, dm, ds, seed=case + 1, n_t=n_t)
        theta = np.stack([np.abs(rng.normal(.1, .05, 5)) + 0.01, rng.normal(10, 3, 5), rng.uniform(0.05, 0.95, 5)], axis=1)
        if kind == "osc_acf_missing":
            for r in range(n_trials):
                s = int(rng.integers(0, n_t - n_t // 10))
                data[r, s:s + n_t // 10] = np.nan
            om = omodels.one_timescale_and_osc_with_missing_model(data, t, prior=po)
            gm = pkg.one_timescale_and_osc_with_missing_model(data, t, "abc", prior=pp)
            tol = 1e-5
        else:
            sm = "acf" if kind == "osc_acf" else "psd"
            om = omodels.one_timescale_and_osc_model(data, t, prior=po, summary_method=sm)
            gm = pkg.one_timescale_and_osc_model(data, t, "abc", prior=pp, summary_method=sm)
            tol = 1e-5
            rel = sm == "psd"
    d_or, s_or = cport.abc_distances(om, theta, seed=4, generation=2, return_stats=True)
    d, s = gm.gpu_model(ctx).distances(theta, seed=4, generation=2, return_stats=True)
    scale = np.max(np.abs(s_or), axis=1, keepdims=True) if rel else 1.0      # PSD: relative to the spectrum's peak
    assert np.max(np.abs(s - s_or) / scale) <= tol
    assert np.all(np.abs(d - d_or) <= 2 * np.sqrt(d_or) * tol + tol ** 2 + 1e-4 * d_or)


def _pgram_cases():
    rng = np.random.default_rng(777)
    return [(case, int(rng.integers(64, 9000)), int(rng.choice([1, 2, 3, 4, 7])), float(rng.choice([100.0, 250.0, 500.0, 1000.0])))
            for case in range(10)]


@pytest.mark.parametrize("case,n_t,n_trials,fs", _pgram_cases())
def test_random_shape_periodogram_model_against_fp64_oracle(pkg, ctx, case, n_t, n_trials, fs):
    """one_timescale_model(summary_method = :psd) at random lengths (nfft = the next 7-smooth integer: zero padding,
    odd radices, odd trial counts) against the NumPy oracle."""
    dt = 1.0 / fs
    t = dt * np.arange(1, n_t + 1)
    data = oou.generate_ou_process(0.05, 1.0, dt, n_t * dt, n_trials, seed=case + 1, n_t=n_t) * 2.0
    lims = (0.5 * fs / n_t * 3, 0.45 * fs)
    om = omodels.one_timescale_model(data, t, summary_method="psd", prior=[omodels.Uniform(0.01, 1.0)], freqlims=lims)
    gm = pkg.one_timescale_model(data, t, "abc", summary_method="psd", prior=[pkg.Uniform(0.01, 1.0)], freqlims=lims)
    assert np.array_equal(om.freq_idx, gm.freq_idx)
    assert np.max(np.abs(gm.data_sum_stats - om.data_sum_stats)) <= 1e-12 * om.data_sum_stats.max()
    theta = np.random.default_rng(50 + case).uniform(0.01, 1.0, size=(3, 1))
    d, s = gm.gpu_model(ctx).distances(theta, seed=6, generation=1, return_stats=True)
    for i in range(3):
        ws = omodels.summary_stats(om, omodels.generate_data(om, theta[i], seed=6, generation=1, particle=i))
        assert np.max(np.abs(s[i] - ws)) <= 1e-5 * ws.max(), (n_t, pkg.summary.nextfastfft(n_t), np.max(np.abs(s[i] - ws)) / ws.max())
        want = omodels.distance_function(om, ws, om.data_sum_stats)
        assert abs(d[i] - want) <= 3e-4 * want


def test_random_spectra_knee_and_peak_against_oracle(pkg, ctx):
    """Random Lorentzian + bump + noise spectra of random lengths / grids through the batched knee kernel."""
    from oracle import knee as oknee
    rng = np.random.default_rng(31)
    for case in range(6):
        n = int(rng.integers(40, 3000))
        f = (np.arange(n) + 1) * float(rng.uniform(0.01, 1.0))
        rows = []
        for r in range(4):
            amp, k = 10 ** rng.uniform(-2, 3), f[int(rng.integers(n // 20 + 1, n // 2))]
            p = amp / (1 + (f / k) ** 2) + 0.2 * amp * np.exp(-(f - f[int(rng.integers(n // 4, n - 2))]) ** 2 / (2 * (3 * f[0]) ** 2))
            rows.append(p * (1 + 0.05 * rng.standard_normal(n)))
        P = np.array(rows)
        got = pkg.utils.find_knee_frequency(P, f, ctx=ctx)
        for r in range(4):
            want = oknee.find_knee_frequency(P[r], f)
            obj = lambda u: float(np.mean(np.abs(oknee.lorentzian(f, u) - P[r])))
            assert obj(got[r]) <= obj(want) * (1 + 1e-9) + 1e-12 * P[r].max()
            assert abs(got[r, 1] - want[1]) <= 1e-5 * want[1], (case, r, got[r], want)
