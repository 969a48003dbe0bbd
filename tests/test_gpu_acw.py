"""GPU parity: acw() (ACW-0 / ACW-50 / ACW-e indices bit-exact, AUC, tau, ACF) vs the oracle."""
import math

import numpy as np
import pytest

from oracle import acw as oacw
from oracle import ou as oou

pytestmark = pytest.mark.gpu

TYPES = ["acw0", "acw50", "acweuler", "auc", "tau"]
ACF_VALUE_TOL = 2e-7


def check_against_oracle(pkg, ctx, data, fs, types=TYPES, n_lags=None, average=False, tau_rtol=1e-5, acf_tol=ACF_VALUE_TOL):
    want = oacw.acw(data, fs, acwtypes=types, n_lags=n_lags, average_over_trials=average)
    got = pkg.acw(data, fs, acwtypes=types, n_lags=n_lags, average_over_trials=average, ctx=ctx)
    res = dict(zip(types, got.acw_results if len(types) > 1 else [got.acw_results]))
    assert got.n_lags == want["n_lags"]
    dt = 1.0 / fs
    for name, key in [("acw0", "k0"), ("acw50", "k50"), ("acweuler", "ke")]:
        if name in types:
            w = np.where(want[key] >= 0, want[key] * dt, np.nan)
            assert np.array_equal(np.atleast_1d(res[name]), w, equal_nan=True), name   # bit-exact
    # values: the streaming pass multiplies in fp32 (accumulates in fp64): ACF error ~1e-8, two
    # orders below the 1e-5 of north_star (1); crossing indices above are exact regardless
    if "auc" in types:
        assert np.allclose(np.atleast_1d(res["auc"]), want["auc"], rtol=1e-6, atol=1e-9)
    if "tau" in types:
        assert np.allclose(np.atleast_1d(res["tau"]), want["tau"], rtol=tau_rtol, atol=0)
    assert got.acf.shape == want["acf"].shape or got.acf.shape == want["acf"][0].shape
    assert np.max(np.abs(np.atleast_2d(got.acf) - want["acf"])) < acf_tol
    assert np.array_equal(got.lags, want["lags"])
    return got, want


def _data_with_k0_first_ge_2(make, seeds=range(100)):
    for s in seeds:
        d = make(s)
        k0 = oacw.acw(d[:1], 1.0, acwtypes=["acw0"])["k0"][0]
        if k0 >= 2:
            return d
    raise RuntimeError("no seed found")


def test_acw_on_white_noise(pkg, ctx):
    d = _data_with_k0_first_ge_2(lambda s: np.random.default_rng(s).standard_normal((400, 5000)))
    # white noise: acf[1] ~ 0, so the exp-decay fit is ill-conditioned (d tau / d acf ~ tau /
    # (acf ln(1/acf))): compare tau loosely and the fitted objective tightly
    got, want = check_against_oracle(pkg, ctx, d, 100.0, tau_rtol=5e-3)
    assert want["n_lags"] < 40                                # short-memory regime (HBM-bound)
    from oracle import acwred
    tau_g = dict(zip(TYPES, got.acw_results))["tau"]
    for i in range(0, 400, 7):
        rg = acwred.expdecay_residual(tau_g[i], want["lags"], want["acf"][i])
        ro = acwred.expdecay_residual(want["tau"][i], want["lags"], want["acf"][i])
        assert rg <= ro * (1 + 1e-9) + 1e-15


def test_acw_on_ou_data(pkg, ctx):
    fs = 100.0
    d = oou.generate_ou_process(0.5, 1.0, 1 / fs, 50.0, 64, seed=3)
    got, want = check_against_oracle(pkg, ctx, d, fs)
    assert want["n_lags"] > 60                                # long-memory regime
    res = dict(zip(TYPES, got.acw_results))
    # reference's own sanity checks (test_acw.jl:128-143, 285-345)
    assert np.nanmean(res["acw50"]) == pytest.approx(0.5 * math.log(2), rel=0.5)
    assert np.mean(res["tau"]) == pytest.approx(0.5, rel=0.5)
    assert np.all(res["auc"] > 0) and np.all(res["auc"] < 50.0)


def test_acw_with_nans_user_lags_vector_and_average(pkg, ctx):
    fs = 100.0
    d = oou.generate_ou_process(0.3, 1.0, 1 / fs, 30.0, 20, seed=4)
    dn = d.copy()
    rng = np.random.default_rng(0)
    for r in range(20):
        s = rng.integers(0, 2500)
        dn[r, s:s + 200] = np.nan
    check_against_oracle(pkg, ctx, dn, fs)
    check_against_oracle(pkg, ctx, dn, fs, n_lags=150)
    check_against_oracle(pkg, ctx, d, fs, n_lags=25, types=["acw0", "acw50", "acweuler", "tau"])  # crossings beyond n_lags
    check_against_oracle(pkg, ctx, d, fs, average=True)
    check_against_oracle(pkg, ctx, d[0], fs)                     # vector input
    one = pkg.acw(d, fs, acwtypes="acw50", ctx=ctx)
    assert one.acwtypes == "acw50" and np.asarray(one.acw_results).shape == (20,)


def test_acw_user_lags_shorter_than_the_zero_crossing(pkg, ctx):
    """Found by benchmarks/acw_fuzz.py.  Long-memory data whose mean ACF has not crossed zero within the
    user's n_lags: (1) average_over_trials used to resume the lag fill at an odd lag count (misaligned
    16-byte loads: the call failed); (2) with NaNs the reference only ever sees n_lags lags, so there is no
    crossing (NaN), while complete data are searched over all lags; (3) the AUC of complete data integrates
    the first slice's whole window even when it is longer than n_lags."""
    rng = np.random.default_rng(5)
    d = np.cumsum(rng.standard_normal((40, 3730)), axis=1) * 0.05 + rng.standard_normal((40, 3730))
    dn = d.copy()
    for r in range(0, 40, 2):
        dn[r, 100 * r:100 * r + 400] = np.nan
    types = ["acw0", "acw50", "acweuler", "tau"]
    for data in (d, dn):
        for n_lags in (455, 100, 33):
            check_against_oracle(pkg, ctx, data, 100.0, types=types, n_lags=n_lags, average=True, tau_rtol=1e-4)
            check_against_oracle(pkg, ctx, data, 100.0, types=types, n_lags=n_lags, tau_rtol=1e-4)
    k0 = oacw.acw(d, 100.0, acwtypes=["acw0"])["k0"]
    assert k0[0] > 40
    got, want = check_against_oracle(pkg, ctx, d, 100.0, n_lags=40, tau_rtol=1e-4)      # AUC window > n_lags
    assert np.all(np.isfinite(dict(zip(TYPES, got.acw_results))["auc"]))


def test_acw_edge_cases(pkg, ctx):
    # never crosses zero -> NaN + n_lags = n (acw.jl:203-207); AUC then throws like the reference
    t = np.arange(600, dtype=np.float64)
    d = np.stack([1.0 + 0.001 * np.sin(0.01 * t), 5.0 + np.cos(0.02 * t)])
    want = oacw.acw(d, 10.0, acwtypes=["acw0", "acw50"])
    got = pkg.acw(d, 10.0, acwtypes=["acw0", "acw50"], ctx=ctx)
    assert np.array_equal(got.acw_results[0], np.where(want["k0"] >= 0, want["k0"] / 10.0, np.nan), equal_nan=True)
    assert got.n_lags == want["n_lags"]
    alt = np.tile(np.array([1.0, -1.0]), 25)                    # first slice crosses zero at lag 1
    with pytest.raises(ValueError):                              # -> 1-point Romberg window: reference throws
        pkg.acw(np.stack([alt, np.arange(50.0)]), 1.0, acwtypes=["auc"], ctx=ctx)
    with pytest.raises(ValueError):
        oacw.acw(np.stack([alt, np.arange(50.0)]), 1.0, acwtypes=["auc"])
    with pytest.raises(pkg.ItjlError):
        pkg.acw(d, -1.0, ctx=ctx)


def test_acw_full_size_c2(pkg, ctx):
    """BASELINE config C2: randn(10000 x 5000), fs = 100, all ACF-based acwtypes.  The oracle
    checks a subset of trials; the rest through size-independent properties."""
    rng = np.random.default_rng(0)
    d = np.asfortranarray(rng.standard_normal((10000, 5000)))
    got = pkg.acw(d, 100.0, acwtypes=["acw0", "acw50", "acweuler", "tau"], ctx=ctx)
    r = dict(zip(["acw0", "acw50", "acweuler", "tau"], got.acw_results))
    idx = np.arange(0, 10000, 97)
    want = oacw.acw(d[idx], 100.0, acwtypes=["acw0", "acw50", "acweuler"], n_lags=got.n_lags)
    for name, key in [("acw0", "k0"), ("acw50", "k50"), ("acweuler", "ke")]:
        assert np.array_equal(r[name][idx], want[key] / 100.0)
    assert np.max(np.abs(got.acf[idx] - want["acf"])) < ACF_VALUE_TOL
    assert np.all(got.acf[:, 0] == 1.0)
    assert np.all(r["acw50"] == 0.01) and np.all(r["acweuler"] == 0.01)     # white noise: lag 1
    assert np.all(r["acw0"] >= 0.01) and got.n_lags == math.ceil(1.1 * round(r["acw0"].max() * 100))
    assert np.all(r["acw0"] >= r["acweuler"]) and np.all(r["acweuler"] >= r["acw50"])


def test_acw_nd_input_and_dims(pkg, ctx):
    """N-d inputs and dims != last (acw.jl:113, get_slices / stack_and_reshape,
    utils.jl:863-895): every slice gives the same numbers as the (slices, time) call, results take
    the shape of the remaining dimensions and the lags sit where time was."""
    rng = np.random.default_rng(5)
    n_t = 1500
    base = np.empty((3, 4, n_t))
    for i in range(3):
        for j in range(4):
            x = np.zeros(n_t); a = np.exp(-1.0 / (5 + 10 * i + 3 * j))
            e = rng.standard_normal(n_t)
            for k in range(1, n_t):
                x[k] = a * x[k - 1] + e[k]
            base[i, j] = x
    types = ["acw0", "acw50", "acweuler", "tau"]
    flat = pkg.acw(base.reshape(12, n_t), 100.0, acwtypes=types, ctx=ctx)          # rows in C order
    r3 = pkg.acw(base, 100.0, acwtypes=types, ctx=ctx)                             # dims = 3 (last)
    moved = np.moveaxis(base, -1, 1)                                               # (3, time, 4)
    rm = pkg.acw(moved, 100.0, acwtypes=types, dims=2, ctx=ctx)
    assert r3.n_lags == flat.n_lags == rm.n_lags
    # ... and against the oracle, slice by slice in the reference's order
    want = oacw.acw(base.reshape(12, n_t), 100.0, acwtypes=types)
    assert r3.n_lags == want["n_lags"]
    for name, key in [("acw0", "k0"), ("acw50", "k50"), ("acweuler", "ke")]:
        w = np.where(want[key] >= 0, want[key] * (1.0 / 100.0), np.nan).reshape(3, 4)
        assert np.array_equal(r3.acw_results[types.index(name)], w, equal_nan=True)
        assert np.array_equal(rm.acw_results[types.index(name)], w, equal_nan=True)
    assert np.max(np.abs(r3.acf.reshape(12, -1) - want["acf"])) < 1e-5
    assert np.allclose(r3.acw_results[3].reshape(-1), want["tau"], rtol=1e-5)
    for k in range(len(types)):
        want = np.asarray(flat.acw_results[k]).reshape(3, 4)
        assert r3.acw_results[k].shape == (3, 4) and rm.acw_results[k].shape == (3, 4)
        assert np.array_equal(r3.acw_results[k], want, equal_nan=True)
        assert np.array_equal(rm.acw_results[k], want, equal_nan=True)
    assert r3.acf.shape == (3, 4, r3.n_lags) and rm.acf.shape == (3, rm.n_lags, 4)
    assert np.array_equal(r3.acf, flat.acf.reshape(3, 4, -1))
    assert np.array_equal(np.moveaxis(rm.acf, 1, -1), r3.acf)
    # 2-D with time first (dims = 1)
    r1 = pkg.acw(base[0].T.copy(), 100.0, acwtypes=types, dims=1, ctx=ctx)
    r2 = pkg.acw(base[0], 100.0, acwtypes=types, ctx=ctx)
    for k in range(len(types)):
        assert np.array_equal(r1.acw_results[k], r2.acw_results[k], equal_nan=True)
    assert r1.acf.shape == (r1.n_lags, 4) and np.array_equal(r1.acf.T, r2.acf)
    with pytest.raises(ValueError):
        pkg.acw(base, 100.0, dims=4, ctx=ctx)


def _random_acw_cases():
    rng = np.random.default_rng(77)
    out = []
    for case in range(12):
        out.append((case, int(rng.choice([1, 2, 7, 33, 129])), int(rng.integers(40, 6000)), float(rng.choice([1.0, 100.0, 500.0])),
                    ["randn", "ou", "ou_offset", "smooth"][int(rng.integers(0, 4))], bool(rng.integers(0, 3) == 0),
                    bool(rng.integers(0, 4) == 0), None if rng.integers(0, 3) else int(rng.integers(2, 700))))
    return out


@pytest.mark.parametrize("case,n_slices,n_t,fs,gen,nan,average,n_lags", _random_acw_cases())
def test_acw_random_inputs_against_oracle(pkg, ctx, case, n_slices, n_t, fs, gen, nan, average, n_lags):
    """A fixed sample of benchmarks/acw_fuzz.py: indices bit-exact, values within the 1e-5 tolerance, the
    same exceptions as the oracle."""
    rng = np.random.default_rng(1000 + case)
    if gen == "randn":
        d = rng.standard_normal((n_slices, n_t))
    elif gen == "smooth":
        d = np.cumsum(rng.standard_normal((n_slices, n_t)), axis=1) * 0.05 + rng.standard_normal((n_slices, n_t))
    else:
        d = oou.generate_ou_process(float(rng.uniform(0.02, 2.0)) * (100.0 / fs), 1.0, 1 / fs, n_t / fs, n_slices, seed=case + 1, n_t=n_t)
        if gen == "ou_offset":
            d = d * float(rng.uniform(0.1, 10)) + float(rng.normal(0, 100))
    if nan:
        for r in range(n_slices):
            if rng.integers(0, 2):
                s = int(rng.integers(0, max(1, n_t - n_t // 8)))
                d[r, s:s + max(1, n_t // 8)] = np.nan
    if n_lags is not None:
        n_lags = min(n_lags, n_t // 2)
    types = ["acw0", "acw50", "acweuler", "tau"]
    try:
        want = oacw.acw(d, fs, acwtypes=types, n_lags=n_lags, average_over_trials=average)
    except ValueError:
        with pytest.raises(ValueError):
            pkg.acw(d, fs, acwtypes=types, n_lags=n_lags, average_over_trials=average, ctx=ctx)
        return
    got = pkg.acw(d, fs, acwtypes=types, n_lags=n_lags, average_over_trials=average, ctx=ctx)
    res = dict(zip(types, got.acw_results))
    assert got.n_lags == want["n_lags"]
    for name, key in [("acw0", "k0"), ("acw50", "k50"), ("acweuler", "ke")]:
        w = np.where(want[key] >= 0, want[key] * (1.0 / fs), np.nan)
        assert np.array_equal(np.atleast_1d(res[name]), w, equal_nan=True), name
    a = np.atleast_2d(got.acf)
    assert a.shape == want["acf"].shape and np.array_equal(np.isnan(a), np.isnan(want["acf"]))
    assert np.nanmax(np.abs(a - want["acf"]), initial=0.0) < 1e-5
    assert np.allclose(np.atleast_1d(res["tau"]), want["tau"], rtol=1e-2, atol=0, equal_nan=True)


def test_acw_skip_zero_lag_three_parameter_fit(pkg, ctx):
    """acw(...; acwtypes = :tau, skip_zero_lag = true) -> fit_expdecay_3_parameters (utils.jl:159-179): A (exp(-t / tau) + B)
    over lags[2:end].  White measurement noise on top of an OU signal puts a delta at lag 0 that the one-parameter fit
    cannot absorb - the case the option exists for."""
    fs, n_t, tau = 100.0, 6000, 0.25
    sig = oou.generate_ou_process(tau, 1.0, 1 / fs, n_t / fs, 9, seed=21, n_t=n_t)
    data = sig + 0.8 * np.random.default_rng(4).standard_normal(sig.shape)
    want = oacw.acw(data, fs, acwtypes=["tau"], skip_zero_lag=True)
    got = pkg.acw(data, fs, acwtypes=["tau"], skip_zero_lag=True, ctx=ctx)
    assert got.n_lags == want["n_lags"]
    assert np.allclose(got.acw_results, want["tau"], rtol=1e-6), (got.acw_results, want["tau"])
    plain = pkg.acw(data, fs, acwtypes=["tau"], ctx=ctx)
    # the three-parameter fit recovers the OU timescale, the one-parameter fit is dragged down by the lag-0 delta
    assert np.all(np.abs(np.log(np.asarray(got.acw_results) / tau)) < np.log(2.2)) and abs(np.median(got.acw_results) / tau - 1.0) < 0.25
    assert np.median(plain.acw_results) < 0.8 * np.median(got.acw_results)
    # a perfect A (exp(-t / tau) + B) row comes back exactly
    lags = np.arange(60) / fs
    row = 0.7 * (np.exp(-lags / 0.11) + 0.05)
    from oracle import acwred
    assert acwred.fit_expdecay_3_parameters(lags, row) == pytest.approx(0.11, rel=1e-8)


@pytest.mark.parametrize("window", [8, 16, 32])
def test_acw_streaming_window_choices(pkg, ctx, window):
    """tuning.acw_stream = 8 / 16 / 32 keeps that many lags in the one-read streaming pass (fewer lags: fewer registers
    and FFMA per sample); slices that have not crossed zero inside the window are redone by the fp64 kernel, so every
    choice must give the oracle's indices bit for bit - on white noise (nothing redone at 32), on an AR(1) mix whose
    zero crossings straddle 8 and 16 lags, and with NaN gaps."""
    rng = np.random.default_rng(11)
    n_slices, n_t = 96, 1500
    white = rng.standard_normal((n_slices, n_t))
    phi = np.linspace(0.0, 0.85, n_slices)[:, None]                     # AR(1): crossings from lag 1 to beyond 16
    ar = np.empty((n_slices, n_t)); prev = rng.standard_normal((n_slices, 1))
    for k in range(n_t):
        prev = phi * prev + np.sqrt(1 - phi ** 2) * rng.standard_normal((n_slices, 1))
        ar[:, k] = prev[:, 0]
    gaps = ar.copy(); gaps[::7, 200:260] = np.nan
    ctx.set_tuning(acw_stream=window)
    try:
        for d in (white, ar, gaps):
            d = d.copy()
            first = next(i for i in range(n_slices) if oacw.acw(d[i], 100.0, acwtypes=["acw0"])["k0"][0] >= 2)
            d[[0, first]] = d[[first, 0]]                               # AUC window of the first slice needs >= 2 points
            check_against_oracle(pkg, ctx, d, 100.0, acf_tol=5e-7)       # fp32 products over 1500 samples
    finally:
        ctx.reset_tuning()


def test_acw_acf_fetch_matches_inline_output(pkg, ctx):
    """C ABI: the ACF matrix through acf_out of itjl_acw (as a C / Julia caller may pass it) and through
    itjl_acw_acf after a call made with acf_out = NULL are the same bytes; the shape query reports (rows, lags)."""
    import ctypes
    lib, L = pkg._lib.lib(), pkg._lib
    rng = np.random.default_rng(5)
    d = np.asfortranarray(rng.standard_normal((300, 900)))
    n_s, n_t = d.shape

    def call(acf_buf):
        k = [np.empty(n_s, dtype=np.int32) for _ in range(3)]
        auc, tau = np.empty(n_s), np.empty(n_s)
        nl = ctypes.c_int64(0)
        ctx.check(lib.itjl_acw(ctx.handle, L.ptr(d), n_s, n_t, 100.0, L.ACW0 | L.ACW50 | L.TAU, 0, ctypes.byref(nl),
                               L.ptr(k[0]), L.ptr(k[1]), L.ptr(k[2]), L.ptr(auc), L.ptr(tau), L.ptr(acf_buf),
                               0 if acf_buf is None else acf_buf.size))
        return int(nl.value), tau
    big = np.full(n_s * n_t, np.nan)
    n_lags, tau_a = call(big)
    n_lags_b, tau_b = call(None)
    assert n_lags == n_lags_b and np.array_equal(tau_a, tau_b)
    rows, lags = ctypes.c_int64(0), ctypes.c_int64(0)
    ctx.check(lib.itjl_acw_acf(ctx.handle, None, 0, ctypes.byref(rows), ctypes.byref(lags)))
    assert (rows.value, lags.value) == (n_s, n_lags)
    exact = np.empty(n_s * n_lags)
    ctx.check(lib.itjl_acw_acf(ctx.handle, L.ptr(exact), exact.size, None, None))
    assert np.array_equal(exact, big[: n_s * n_lags]) and np.all(np.isnan(big[n_s * n_lags:]))
    assert np.allclose(exact.reshape((n_s, n_lags), order="F")[:, 0], 1.0)
    small = np.empty(n_s * n_lags - 1)
    assert lib.itjl_acw_acf(ctx.handle, L.ptr(small), small.size, None, None) != 0          # capacity too small: refused


def test_acw_many_slices_overflow_the_pinned_scratch(pkg, ctx):
    """170 000 short slices: the crossing indices (3 x 680 KB) no longer fit the 2 MB pinned scratch the small
    device -> host results are staged through, so part of them takes the direct copy - every index must still be
    the oracle's, and the fetched ACF must be the right shape."""
    rng = np.random.default_rng(8)
    n_slices, n_t = 170_000, 300
    phi = rng.uniform(0.0, 0.6, (n_slices, 1))
    d = np.empty((n_slices, n_t)); prev = rng.standard_normal((n_slices, 1))
    for k in range(n_t):
        prev = phi * prev + np.sqrt(1 - phi ** 2) * rng.standard_normal((n_slices, 1))
        d[:, k] = prev[:, 0]
    types = ["acw0", "acw50", "acweuler"]
    want = oacw.acw(d, 50.0, acwtypes=types)
    got = pkg.acw(d, 50.0, acwtypes=types, ctx=ctx)
    assert got.n_lags == want["n_lags"] and got.acf.shape == (n_slices, want["n_lags"])
    dt = 1.0 / 50.0
    for name, key, res in zip(types, ("k0", "k50", "ke"), got.acw_results):
        w = np.where(want[key] >= 0, want[key] * dt, np.nan)
        assert np.array_equal(res, w, equal_nan=True), name
    assert np.max(np.abs(got.acf - want["acf"])) < 5e-6      # fp32 products of the streaming pass, max over 1.7e7 values (north_star: 1e-5)
