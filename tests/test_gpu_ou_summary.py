"""GPU parity: OU generators and summary statistics on stored data vs the fp64 oracle.
All calls go through the C ABI (ctypes)."""
import numpy as np
import pytest

from oracle import ou as oou
from oracle import summary as osum

pytestmark = pytest.mark.gpu

TRAJ_TOL = 1e-5      # north_star (2): trajectories within 1e-5 (fp32 recurrence vs fp64)


@pytest.mark.parametrize("update", ["exact", "em"])
@pytest.mark.parametrize("standardize", [True, False])
@pytest.mark.parametrize("n_t,n_trials,tau", [(5000, 2, 0.3), (1001, 5, 0.05), (4097, 3, 2.0)])
def test_generate_ou_injected_noise(pkg, ctx, update, standardize, n_t, n_trials, tau):
    rng = np.random.default_rng(n_t)
    dt = 0.002
    u0 = rng.standard_normal(n_trials)
    noise = rng.standard_normal((n_trials, n_t))
    want = oou.generate_ou_process(tau, 1.7, dt, n_t * dt, n_trials, standardize, u0=u0, noise=noise,
                                   update=update, n_t=n_t)
    got = pkg.generate_ou_process(tau, 1.7, dt, n_t * dt, n_trials, standardize, update=update,
                                  u0=u0, noise=noise.astype(np.float64), ctx=ctx)
    assert got.shape == (n_trials, n_t) and got.flags.f_contiguous
    scale = np.abs(want).max()
    assert np.max(np.abs(got - want)) <= TRAJ_TOL * scale


def test_generate_ou_philox_stream_matches_oracle(pkg, ctx):
    dt, n_t, n_trials = 0.002, 5000, 4
    want = oou.generate_ou_process(0.3, 1.0, dt, n_t * dt, n_trials, seed=77, n_t=n_t)
    got = pkg.generate_ou_process(0.3, 1.0, dt, n_t * dt, n_trials, seed=77, ctx=ctx)
    assert np.max(np.abs(got - want)) <= TRAJ_TOL * np.abs(want).max()
    # reference's statistical checks (test_ou_process.jl:18-31)
    assert np.allclose(got.std(axis=1, ddof=1), 1.0, atol=1e-5)
    ac = np.mean([np.corrcoef(r[:-1], r[1:])[0, 1] for r in got])
    assert abs(ac - np.exp(-dt / 0.3)) < 0.01
    again = pkg.generate_ou_process(0.3, 1.0, dt, n_t * dt, n_trials, seed=77, ctx=ctx)
    assert np.array_equal(got, again)                       # same seed -> identical
    other = pkg.generate_ou_process(0.3, 1.0, dt, n_t * dt, n_trials, seed=78, ctx=ctx)
    assert not np.allclose(got, other)


def test_generate_ou_nan_contract_and_args(pkg, ctx):
    bad = pkg.generate_ou_process(-1e-4, 1.0, 0.002, 4.0, 2, seed=1, ctx=ctx)   # a = exp(+20) per step
    assert np.isnan(bad).all()
    with pytest.raises(pkg.ItjlError):
        pkg.generate_ou_process(0.3, 1.0, 0.002, 0.004, 2, ctx=ctx)      # n_t = 2 < 4


@pytest.mark.parametrize("coeff", [0.95, 0.3, 1.4, -0.2])
def test_generate_ou_with_oscillation(pkg, ctx, coeff):
    dt, n_t, n_trials = 1 / 500, 10000, 3
    rng = np.random.default_rng(3)
    u0 = rng.standard_normal(n_trials); ph = rng.random(n_trials) * 2 * np.pi
    noise = rng.standard_normal((n_trials, n_t))
    th = [0.05, 10.0, coeff]
    want = oou.generate_ou_with_oscillation(th, dt, n_t * dt, n_trials, 0.7, 2.5, u0=u0, noise=noise,
                                            phases=ph, n_t=n_t)
    got = pkg.generate_ou_with_oscillation(th, dt, n_t * dt, n_trials, 0.7, 2.5, u0=u0, noise=noise,
                                           phases=ph, ctx=ctx)
    assert np.max(np.abs(got - want)) <= 1e-5 * np.abs(want).max()
    want2 = oou.generate_ou_with_oscillation(th, dt, n_t * dt, n_trials, 0.7, 2.5, seed=9, n_t=n_t)
    got2 = pkg.generate_ou_with_oscillation(th, dt, n_t * dt, n_trials, 0.7, 2.5, seed=9, ctx=ctx)
    assert np.max(np.abs(got2 - want2)) <= 1e-5 * np.abs(want2).max()


# --------------------------------------------------------------- summary stats ----
def _ou_data(n_trials, n_t, tau=0.3, seed=1):
    return oou.generate_ou_process(tau, 1.0, 0.002, n_t * 0.002, n_trials, seed=seed, n_t=n_t)


@pytest.mark.parametrize("n_trials,n_t,n_lags", [(7, 5000, 480), (3, 1000, 1000), (1, 333, 20), (40, 2048, 1)])
def test_comp_ac_fft_matches_oracle(pkg, ctx, n_trials, n_t, n_lags):
    x = _ou_data(n_trials, n_t)
    want = osum.comp_ac_fft(x, n_lags)
    got = pkg.comp_ac_fft(x, n_lags, ctx=ctx)
    assert got.shape == want.shape
    assert np.max(np.abs(got - want)) < 1e-12               # fp64 direct sum vs fp64 FFT
    assert np.all(got[:, 0] == 1.0)
    v = pkg.comp_ac_fft(x[0], n_lags, ctx=ctx)
    assert v.shape == (n_lags,) and np.max(np.abs(v - want[0])) < 1e-12


def test_comp_ac_time_missing_matches_oracle(pkg, ctx):
    x = _ou_data(6, 3000, seed=4)
    rng = np.random.default_rng(0)
    for r in range(6):
        s = rng.integers(0, 2500)
        x[r, s:s + 300] = np.nan
    x[2, rng.integers(0, 3000, 200)] = np.nan
    want = osum.comp_ac_time_missing(x, 300)
    got = pkg.comp_ac_time_missing(x, 300, ctx=ctx)
    assert np.max(np.abs(got - want)) < 1e-12
    # complete data: identical to comp_ac_fft (test_summary_stats.jl:57-74)
    y = _ou_data(4, 2000)
    assert np.max(np.abs(pkg.comp_ac_time_missing(y, 100, ctx=ctx) - pkg.comp_ac_fft(y, 100, ctx=ctx))) < 1e-13


@pytest.mark.parametrize("n_trials,n_t", [(3, 10000), (5, 1000), (2, 4096), (2, 5000)])
def test_comp_psd_adfriendly_matches_oracle(pkg, ctx, n_trials, n_t):
    dt = 1 / 500
    x = oou.generate_ou_with_oscillation([0.05, 10.0, 0.9], dt, n_t * dt, n_trials, 0.3, 1.5, seed=2, n_t=n_t)
    want, wf = osum.comp_psd_adfriendly(x, 1 / dt)
    got, gf = pkg.comp_psd_adfriendly(x, 1 / dt, ctx=ctx)
    assert got.shape == want.shape and np.allclose(gf, wf, rtol=1e-14)
    # fp32 FFT: error is relative to the spectrum's peak, 1e-5 relative on the bins that
    # carry the statistic (f < 100 Hz)
    assert np.max(np.abs(got - want)) <= 1e-5 * want.max()
    sel = (wf > 0.5) & (wf < 100.0)
    rel = np.abs(got[:, sel] - want[:, sel]) / want[:, sel]
    assert np.median(rel) < 1e-5 and rel.max() < 1e-3


def test_summary_argument_errors(pkg, ctx):
    with pytest.raises(pkg.ItjlError):
        pkg.comp_ac_fft(np.zeros((2, 100)), 101, ctx=ctx)
    with pytest.raises(pkg.ItjlError):
        pkg.comp_psd_adfriendly(np.zeros((2, 100)), -1.0, ctx=ctx)


def test_two_timescale_generator(pkg, ctx):
    """generate_data of the reference's TwoTimescaleModel (src/models/two_timescale.jl:30-51)."""
    theta, dt, n_t, n_trials = [0.02, 0.4, 0.3], 0.002, 4000, 5
    want = oou.generate_two_timescale(theta, dt, n_t * dt, n_trials, data_var=2.5, seed=9, n_t=n_t)
    got = pkg.generate_two_timescale(theta, dt, n_t * dt, n_trials, data_var=2.5, seed=9, ctx=ctx)
    assert got.shape == want.shape and np.max(np.abs(got - want)) <= 1e-5 * np.abs(want).max()
    # each component is standardised before mixing: var = data_var * (coeff / tau1^2 + (1 - coeff) / tau2^2) up to the
    # sample cross-correlation of the two independent processes
    v = 2.5 * (0.3 / 0.02 ** 2 + 0.7 / 0.4 ** 2)
    assert abs(got.var(axis=1, ddof=1).mean() / v - 1.0) < 0.1
    with pytest.raises(pkg.ItjlError):
        pkg.generate_two_timescale([0.02, 0.4, 1.3], dt, n_t * dt, n_trials, ctx=ctx)
