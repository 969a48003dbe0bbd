"""GPU parity of the DSP-convention periodogram path: comp_psd on stored data, the fused
simulate -> periodogram -> distance kernel behind one_timescale_model(summary_method = :psd), and the
Lomb-Scargle periodogram of data with gaps - each against the fp64 oracle, through the C ABI."""
import numpy as np
import pytest

from oracle import models as omodels
from oracle import ou as oou
from oracle import summary as osum

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n_slices,n_t,fs", [(7, 5000, 500.0), (3, 1000, 100.0), (4, 4999, 250.0), (5, 37, 10.0), (1, 2401, 1.0),
                                             (2, 11000, 1000.0)])
def test_comp_psd_periodogram_on_stored_data(pkg, ctx, n_slices, n_t, fs):
    rng = np.random.default_rng(n_t)
    data = rng.standard_normal((n_slices, n_t)).cumsum(axis=1) * 0.05 + rng.standard_normal((n_slices, n_t)) + 0.3
    want, wf = osum.comp_psd_periodogram(data, fs)
    got, gf = pkg.summary.comp_psd(data, fs, ctx=ctx)
    assert got.shape == want.shape and np.allclose(gf, wf, rtol=1e-14, atol=0.0)
    assert pkg.summary.nextfastfft(n_t) == osum.nextfastfft(n_t)
    # fp64 transform: relative to the largest bin of the slice (what an FFT's rounding error scales with)
    err = np.abs(got - want) / want.max(axis=1, keepdims=True)
    assert err.max() < 1e-12, err.max()
    v = pkg.summary.comp_psd(data[0], fs, ctx=ctx)[0]
    assert v.shape == (want.shape[1],) and np.max(np.abs(v - got[0])) < 1e-12 * want[0].max()    # (paired with another slice above)


def _make(pkg, n_t, n_trials, dt, tau=0.05, seed=3, freqlims=None, distance_method=None):
    t = dt * np.arange(1, n_t + 1)
    data = oou.generate_ou_process(tau, 1.0, dt, n_t * dt, n_trials, seed=seed, n_t=n_t) * 1.7 + 0.4
    om = omodels.one_timescale_model(data, t, summary_method="psd", prior=[omodels.Uniform(0.01, 1.0)],
                                     freqlims=freqlims, distance_method=distance_method)
    gm = pkg.one_timescale_model(data, t, "abc", summary_method="psd", prior=[pkg.Uniform(0.01, 1.0)],
                                 freqlims=freqlims, distance_method=distance_method)
    return om, gm


@pytest.mark.parametrize("n_t,n_trials,dt,freqlims,dist", [(5000, 2, 0.002, None, None), (5000, 5, 0.002, (1.0, 250.1), "linear"),
                                                           (2401, 3, 0.004, (0.2, 60.0), None), (1234, 4, 0.01, None, None)])
def test_psd_model_and_fused_distances(pkg, ctx, n_t, n_trials, dt, freqlims, dist):
    """one_timescale_model(:psd): constructor fields, then distances and trial-mean spectra of the fused kernel with
    injected noise against the oracle's generate_data -> comp_psd -> distance."""
    om, gm = _make(pkg, n_t, n_trials, dt, freqlims=freqlims, distance_method=dist)
    assert gm.summary_method == "psd" and gm.distance_method == om.distance_method
    assert np.array_equal(om.freq_idx, gm.freq_idx) and np.allclose(om.lags_freqs, gm.lags_freqs, rtol=1e-14)
    assert np.max(np.abs(gm.data_sum_stats - om.data_sum_stats) / om.data_sum_stats.max()) < 1e-12
    g = gm.gpu_model(ctx)
    assert g.kernel_info()[0] == "k_abc_pgram"
    rng = np.random.default_rng(n_t + n_trials)
    theta = np.array([[0.05], [0.011], [0.3], [0.9]])
    P = theta.shape[0]
    u0 = rng.standard_normal((P, n_trials)); noise = rng.standard_normal((P, n_trials, n_t))
    got, stats = g.distances(theta, u0=u0, noise=noise, return_stats=True)
    for i in range(P):
        x = omodels.generate_data(om, theta[i], u0=u0[i], noise=noise[i])
        want_stats = omodels.summary_stats(om, x)
        want = omodels.distance_function(om, want_stats, om.data_sum_stats)
        # fp32 transform of an fp32 trajectory: 1e-5 of the trial-mean spectrum's largest selected bin
        assert np.max(np.abs(stats[i] - want_stats)) <= 1e-5 * want_stats.max(), (i, np.max(np.abs(stats[i] - want_stats)) / want_stats.max())
        rel = np.abs(stats[i] - want_stats) / want_stats
        if om.distance_method == "logarithmic":
            # d = mean(log-ratio^2): bound from the per-bin relative errors actually seen
            lr = np.log(want_stats) - np.log(om.data_sum_stats)
            bound = 2.0 * np.mean(np.abs(lr) * rel) + np.mean(rel ** 2) + 1e-9 * want
            assert abs(got[i] - want) <= 1.5 * bound + 1e-12, (got[i], want, bound)
            assert abs(got[i] - want) <= 2e-4 * want
        else:
            assert abs(got[i] - want) <= 2e-5 * want + 1e-12 * want_stats.max() ** 2, (got[i], want)
    # Philox mode: batching invariance and determinism
    th = rng.uniform(0.02, 0.8, size=(24, 1))
    a = g.distances(th, seed=9, generation=1)
    b = np.concatenate([g.distances(th[:7], seed=9, generation=1), g.distances(th[7:], seed=9, generation=1, particle_offset=7)])
    assert np.array_equal(a, b) and np.all(np.isfinite(a))
    # odd trial counts leave the imaginary half of the last transform empty: mean over trials must not change meaning
    d, isacc, n_total, n_acc = g.generation(th, float(np.median(a)), 5, seed=9, generation=1)
    assert np.array_equal(d[:n_total], a[:n_total]) and n_acc == 5


def test_psd_model_philox_against_oracle(pkg, ctx):
    om, gm = _make(pkg, 3000, 3, 0.002, seed=8)
    theta = np.array([[0.03], [0.1], [0.6]])
    got, stats = gm.gpu_model(ctx).distances(theta, seed=4, generation=2, return_stats=True)
    for i in range(theta.shape[0]):
        x = omodels.generate_data(om, theta[i], seed=4, generation=2, particle=i)
        ws = omodels.summary_stats(om, x)
        assert np.max(np.abs(stats[i] - ws)) <= 1e-5 * ws.max()
        assert abs(got[i] - omodels.distance_function(om, ws, om.data_sum_stats)) <= 2e-4 * got[i]


@pytest.mark.parametrize("n_slices,n_t,dt", [(3, 400, 0.01), (4, 1501, 0.002), (2, 5000, 0.002)])
def test_lombscargle_on_data_with_gaps(pkg, ctx, n_slices, n_t, dt):
    rng = np.random.default_rng(n_t)
    t = dt * np.arange(1, n_t + 1)
    data = np.sin(2 * np.pi * 3.3 * t)[None, :] + 0.5 * rng.standard_normal((n_slices, n_t)) + 0.7
    mask = np.zeros((n_slices, n_t), dtype=bool)
    for r in range(n_slices):
        s = rng.integers(0, n_t - n_t // 8)
        mask[r, s:s + n_t // 10] = True
        mask[r, rng.integers(0, n_t, 20)] = True
    data[mask] = np.nan
    want, wf = osum.comp_psd_lombscargle(t, data, mask, dt)
    got, gf = pkg.summary.comp_psd_lombscargle(t, data, mask, dt, ctx=ctx)
    assert got.shape == want.shape and np.allclose(gf, wf, rtol=1e-12, atol=0.0)
    # the last grid point of an even-length record is the Nyquist frequency, where sin(omega t) vanishes on the grid:
    # the one-basis-function limit there, and finite everywhere
    assert np.all(np.isfinite(got))
    assert np.max(np.abs(got - want)) < 1e-9
    assert np.all(got >= -1e-12) and np.all(got <= 1.0 + 1e-12)


@pytest.mark.parametrize("n_slices,n_t,dt", [(5, 1000, 0.01), (3, 777, 0.002), (4, 5000, 0.01), (3, 18000, 0.001)])
def test_lombscargle_chirp_z_matches_direct_sums(pkg, ctx, n_slices, n_t, dt):
    """The default path evaluates the Lomb-Scargle sums as chirp-z transforms (k_lomb_czt, O(M log M) per slice);
    tuning.psd_v2 bit 1 forces the direct O(n_t n_freq) sums (k_lomb_scargle): both must agree to rounding, and with the
    oracle, on even / odd / long records, with gaps at the edges and a slice that is nearly all missing."""
    rng = np.random.default_rng(n_t + 1)
    t = dt * np.arange(1, n_t + 1)
    data = np.sin(2 * np.pi * (0.07 / dt) * t)[None, :] + rng.standard_normal((n_slices, n_t)).cumsum(axis=1) * 0.05 + 3.0
    data[0, : n_t // 7] = np.nan                       # gap at the start
    data[1, -n_t // 9:] = np.nan                       # gap at the end
    data[2, rng.random(n_t) < 0.8] = np.nan            # mostly missing
    mask = np.isnan(data)
    n0 = ctx.launch_count
    got, gf = pkg.comp_psd_lombscargle(t, data, mask, dt, ctx=ctx)
    assert ctx.launch_count - n0 == 2                  # prepare + transform kernel (no relayout: slice-fastest input)
    with ctx.tuned(psd_v2=3):
        direct, df_ = pkg.comp_psd_lombscargle(t, data, mask, dt, ctx=ctx)
    assert np.array_equal(gf, df_) and got.shape == direct.shape
    print(f"n_t={n_t}: chirp-z vs direct max abs diff {np.max(np.abs(got - direct)):.2e}")
    assert np.max(np.abs(got - direct)) < 1e-10
    if n_t <= 5000:
        want, _ = osum.comp_psd_lombscargle(t, data, mask, dt)
        assert np.max(np.abs(got - want)) < 1e-9
