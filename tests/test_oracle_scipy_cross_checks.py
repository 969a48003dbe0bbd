"""Independent cross-checks of the oracle's restatements of THIRD-PARTY conventions (DSP.jl periodogram, Romberg.jl,
LombScargle.jl, StatsBase.percentile, NonlinearSolve's exp-decay fit) against SciPy / NumPy implementations of the
same published definitions.  This does not pin them to the Julia packages (no Julia here: PARITY UNPINNED stands,
tests/test_reference_golden.py is the path to that) - it shows that the restated conventions are the standard ones
an independent implementation arrives at.  CPU only."""
import math

import numpy as np
import pytest
from scipy import integrate, optimize, signal

from oracle import abc as oabc
from oracle import acwred
from oracle import summary as osum


@pytest.mark.parametrize("n,fs", [(5000, 500.0), (4999, 250.0), (1000, 100.0), (37, 10.0), (2401, 1.0)])
def test_periodogram_conventions_against_scipy(n, fs):
    """DSP.periodogram(x; fs, window = hamming): no detrending, SYMMETRIC Hamming window of the signal's length, zero
    padding to nfft, one-sided density |X|^2 / (fs sum w^2) doubled except at DC and Nyquist - scipy.signal.periodogram
    with an explicit window array, detrend = False, scaling = 'density' is the same estimator."""
    rng = np.random.default_rng(n)
    x = rng.standard_normal(n).cumsum() * 0.1 + rng.standard_normal(n) + 0.7
    nfft = osum.nextfastfft(n)
    w = signal.get_window("hamming", n, fftbins=False)
    assert np.allclose(w, osum.hamming_window(n), rtol=0, atol=1e-15)
    f, p = signal.periodogram(x, fs=fs, window=w, nfft=nfft, detrend=False, return_onesided=True, scaling="density")
    po, fo = osum.comp_psd_periodogram_vec(x, fs)
    assert np.allclose(fo, f[1:], rtol=1e-13) and np.allclose(po, p[1:], rtol=1e-10, atol=1e-14 * p.max())
    # 7-smooth padding length
    m = nfft
    for q in (2, 3, 5, 7):
        while m % q == 0:
            m //= q
    assert m == 1 and nfft >= n and all(osum.nextfastfft(k) == k for k in (1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 5000, 16384))


def test_adfriendly_psd_against_scipy():
    """comp_psd_adfriendly (summary.jl:204-235): demeaned, Hamming, zero padded to 2 nextpow2(n), |X|^2 / (fs sum w^2)
    WITHOUT the one-sided doubling = half of SciPy's one-sided density away from DC."""
    rng = np.random.default_rng(3)
    n, fs = 3000, 500.0
    x = rng.standard_normal(n).cumsum() * 0.05 + rng.standard_normal(n)
    n2 = 2 * 4096
    w = signal.get_window("hamming", n, fftbins=False)
    f, p = signal.periodogram(x, fs=fs, window=w, nfft=n2, detrend="constant", return_onesided=True, scaling="density")
    po, fo = osum.comp_psd_adfriendly_vec(x, fs)
    assert np.allclose(fo, f[1:n2 // 2], rtol=1e-13) and np.allclose(po, 0.5 * p[1:n2 // 2], rtol=1e-9, atol=1e-14 * p.max())


@pytest.mark.parametrize("k", [1, 2, 3, 5, 8])
def test_romberg_against_scipy_on_dyadic_grids(k):
    """Romberg.jl on 2^k + 1 equally spaced samples is the classic Romberg table (scipy.integrate.romb)."""
    n = 2 ** k + 1
    x = np.linspace(0.0, 2.0, n)
    for y in (np.exp(-x), np.cos(3 * x) + 0.2 * x ** 3, 1.0 / (1.0 + x * x)):
        want = integrate.romb(y, dx=x[1] - x[0])
        got, err = acwred.romberg(x[1] - x[0], y)
        # Richardson.jl does not always run the full table: it stops once its error estimate is below sqrt(eps) of the
        # value, or when a column stops converging (breaktol) - the two agree within the estimate it reports
        assert abs(got - want) <= max(4.0 * err, 2e-8 * abs(want)), (k, got, want, err)


def test_romberg_on_other_lengths_is_a_richardson_table_of_trapezoid_sums():
    """For n - 1 with odd factors Romberg.jl extrapolates the trapezoid sums at the strides given by the prime
    factorisation: exact for polynomials up to the order the table reaches, and at least as good as Simpson on smooth
    integrands."""
    for n in (7, 10, 13, 16, 31, 101):
        x = np.linspace(0.0, 1.0, n)
        assert acwred.acw_romberg(x[1] - x[0], 3 * x ** 2 + 1) == pytest.approx(2.0, rel=1e-12)
        got = acwred.acw_romberg(x[1] - x[0], np.exp(x))
        assert abs(got - (math.e - 1)) <= abs(integrate.trapezoid(np.exp(x), dx=x[1] - x[0]) - (math.e - 1))


def test_lombscargle_against_scipy():
    """Generalised (floating-mean) Lomb-Scargle, standard normalisation = scipy.signal.lombscargle(floating_mean = True,
    normalize = True) on mean-centred data (away from the degenerate Nyquist point of a regular grid)."""
    rng = np.random.default_rng(0)
    n, dt = 600, 0.01
    t = dt * np.arange(1, n + 1)
    y = np.sin(2 * np.pi * 3.3 * t) + 0.5 * rng.standard_normal(n) + 0.7
    keep = np.ones(n, bool); keep[100:190] = False; keep[rng.integers(0, n, 30)] = False
    f = osum.lombscargle_autofrequency(t, 1 / (2 * dt))
    assert f[0] == pytest.approx(0.5 * (f[1] - f[0])) and (f[1] - f[0]) == pytest.approx(1.0 / (5 * (t[-1] - t[0])))
    p = osum.lombscargle_vec(t[keep], y[keep], f)
    try:
        q = signal.lombscargle(t[keep], y[keep] - y[keep].mean(), 2 * np.pi * f, floating_mean=True, normalize=True)
    except TypeError:
        pytest.skip("this SciPy has no floating_mean")
    ok = np.abs(f - 0.5 / dt) > 1e-9
    assert np.max(np.abs(p[ok] - q[ok])) < 1e-12


def test_percentile_is_quantile_type_7():
    """StatsBase.percentile = Hyndman-Fan type 7 = NumPy's default 'linear' method (abc.jl:729-731)."""
    rng = np.random.default_rng(5)
    x = rng.gamma(2.0, 1.0, 1001)
    for p in (0.0, 25.0, 50.0, 73.3, 100.0):
        s = np.sort(x); pos = p / 100 * (x.size - 1); lo = int(math.floor(pos))
        want = s[lo] + (s[min(lo + 1, x.size - 1)] - s[lo]) * (pos - lo)
        assert oabc.percentile_type7(x, p) == pytest.approx(want, rel=1e-14)


def test_fit_expdecay_is_the_least_squares_argmin():
    """fit_expdecay (utils.jl:113-125: LevenbergMarquardt on exp(-lags / tau) - acf) returns the least-squares argmin:
    the same point SciPy's bounded scalar minimiser and its LM (least_squares, method = 'lm') reach."""
    rng = np.random.default_rng(2)
    lags = 0.002 * np.arange(400)
    for tau in (0.02, 0.1, 0.3):
        acf = np.exp(-lags / tau) + 0.02 * rng.standard_normal(lags.size)
        got = acwred.fit_expdecay(lags, acf)
        r = optimize.least_squares(lambda u: np.exp(-lags / u[0]) - acf, [tau * 1.7], method="lm", xtol=1e-14, ftol=1e-14)
        assert got == pytest.approx(r.x[0], rel=1e-7)
