"""GPU parity of the warp-specialised PSD kernel k_abc_psd2 (4096 < n_t <= 16384: pruned first stage, two
4096-point sub-transforms side by side, simulate under the transform) against the fp64 oracle and against the
first kernel k_abc_psd (tuning.psd_v2 = 0) on the same Philox streams."""
import numpy as np
import pytest

from oracle import models as omodels
from oracle import ou as oou

pytestmark = pytest.mark.gpu


def _models(pkg, n_t, n_trials, dt, osc, freqlims, seed=11):
    t = dt * np.arange(1, n_t + 1)
    if osc:
        data = oou.generate_ou_with_oscillation([0.05, 10.0, 0.9], dt, n_t * dt, n_trials, 0.3, 1.4, seed=seed, n_t=n_t)
        po = [omodels.Normal(.1, .1), omodels.Normal(10, 5), omodels.Uniform(0, 1)]
        pp = [pkg.Normal(.1, .1), pkg.Normal(10, 5), pkg.Uniform(0, 1)]
    else:
        data = oou.generate_ou_process(0.05, 1.0, dt, n_t * dt, n_trials, seed=seed, n_t=n_t)
        po = [omodels.Normal(.1, .1), omodels.Normal(10, 5), omodels.Uniform(0, 1)]
        pp = [pkg.Normal(.1, .1), pkg.Normal(10, 5), pkg.Uniform(0, 1)]
    om = omodels.one_timescale_and_osc_model(data, t, prior=po, summary_method="psd", freqlims=freqlims)
    gm = pkg.one_timescale_and_osc_model(data, t, "abc", prior=pp, summary_method="psd", freqlims=freqlims)
    return om, gm


@pytest.mark.parametrize("n_t,n_trials,dt,freqlims", [
    (10000, 3, 1 / 500, None),             # C4's trial shape: N = 16384, four sub-transforms in two rounds
    (5000, 4, 1 / 500, None),              # N = 8192: two sub-transforms, one round
    (8193, 2, 1 / 500, (0.2, 240.0)),      # smallest N = 16384 shape; one sample beyond the first quarter
    (16384, 1, 1 / 1000, (1.0, 80.0)),     # largest trial: the packed data fill exactly half of the transform
    (4097, 5, 1 / 250, (0.3, 124.0)),      # smallest N = 8192 shape, odd length (last packed pair half empty)
    (8192, 2, 1 / 500, None),
    (12345, 2, 1 / 500, (0.5, 30.0)),
])
def test_psd2_against_oracle_and_first_kernel(pkg, ctx, n_t, n_trials, dt, freqlims):
    om, gm = _models(pkg, n_t, n_trials, dt, True, freqlims)
    g2 = gm.gpu_model(ctx)
    assert g2.kernel_info()[0] == "k_abc_psd2"
    rng = np.random.default_rng(n_t)
    theta = np.array([[0.05, 10.0, 0.9], [0.1, 12.0, 0.5], [0.02, 8.0, 1.2], [0.5, 30.0, 0.05]])
    P = theta.shape[0]
    u0 = rng.standard_normal((P, n_trials)); noise = rng.standard_normal((P, n_trials, n_t))
    phases = rng.uniform(0, 2 * np.pi, (P, n_trials))
    got, stats = g2.distances(theta, u0=u0, noise=noise, phases=phases, return_stats=True)
    for i in range(P):
        x = omodels.generate_data(om, theta[i], u0=u0[i], noise=noise[i], phases=phases[i])
        ws = omodels.summary_stats(om, x)
        want = omodels.distance_function(om, ws, om.data_sum_stats)
        # fp32 transform: 1e-5 of the largest selected bin of the trial-mean spectrum
        assert np.max(np.abs(stats[i] - ws)) <= 1e-5 * ws.max(), (i, np.max(np.abs(stats[i] - ws)) / ws.max())
        assert abs(got[i] - want) <= 3e-4 * want + 1e-12, (i, got[i], want)
    # same Philox streams through the first kernel
    th = rng.uniform([0.02, 5.0, 0.0], [0.5, 40.0, 1.0], size=(16, 3))
    a, sa = g2.distances(th, seed=3, generation=1, return_stats=True)
    with ctx.tuned(psd_v2=0):
        gm1 = pkg.one_timescale_and_osc_model(om.data, om.time, "abc", prior=gm.prior, summary_method="psd", freqlims=freqlims)
        g1 = gm1.gpu_model(ctx)
        assert g1.kernel_info()[0] == "k_abc_psd"
        b, sb = g1.distances(th, seed=3, generation=1, return_stats=True)
    assert np.max(np.abs(sa - sb) / sb.max(axis=1, keepdims=True)) < 5e-6
    assert np.all(np.abs(a - b) <= 3e-4 * b)
    # batching invariance / determinism
    c = np.concatenate([g2.distances(th[:5], seed=3, generation=1), g2.distances(th[5:], seed=3, generation=1, particle_offset=5)])
    assert np.array_equal(a, c)


def test_psd2_plain_ou_kind_and_many_trials(pkg, ctx):
    """kind = OU with the Hamming PSD summary (no oscillation term) and an odd number of trials: the two trial buffers
    alternate through full / empty barriers 51 times."""
    import ctypes
    n_t, n_trials, dt = 9000, 51, 1 / 500
    t = dt * np.arange(1, n_t + 1)
    data = oou.generate_ou_process(0.05, 1.0, dt, n_t * dt, 3, seed=2, n_t=n_t)
    gm = pkg.one_timescale_and_osc_model(data, t, "abc", prior=[pkg.Uniform(0, 1)] * 3, summary_method="psd")
    om = omodels.one_timescale_and_osc_model(data, t, prior=[omodels.Uniform(0, 1)] * 3, summary_method="psd")
    gm.numTrials = om.numTrials = n_trials
    theta = np.array([[0.05, 10.0, 0.95], [0.2, 20.0, 0.3]])
    got, stats = gm.gpu_model(ctx).distances(theta, seed=1, generation=0, return_stats=True)
    for i in range(2):
        x = omodels.generate_data(om, theta[i], seed=1, generation=0, particle=i)
        ws = omodels.summary_stats(om, x)
        assert np.max(np.abs(stats[i] - ws)) <= 1e-5 * ws.max()
        assert abs(got[i] - omodels.distance_function(om, ws, om.data_sum_stats)) <= 3e-4 * got[i]


@pytest.mark.parametrize("n_t,hi", [(15642, 0.49), (16384, 0.49), (9000, 0.49)])
def test_large_trials_with_a_wide_band_fit_or_fall_back(pkg, ctx, n_t, hi):
    """Shapes at the shared-memory limit (found by benchmarks/psd_fuzz.py: n_t = 15642 with 16 040 selected bins was
    accepted by itjl_model_create and then refused by the launch - the static shared memory and the per-CTA reserve were
    not counted): the model either runs k_abc_psd2 or falls back to k_abc_psd with its bin sums in global memory, and a
    refused launch must not poison the next call."""
    from oracle import cport
    fs = 1000.0
    dt = 1 / fs
    t = dt * np.arange(1, n_t + 1)
    data = oou.generate_ou_with_oscillation([0.05, 10.0, 0.9], dt, n_t * dt, 2, 0.0, 1.0, seed=1, n_t=n_t)
    po = [omodels.Normal(.1, .1), omodels.Normal(10, 5), omodels.Uniform(0, 1)]
    pp = [pkg.Normal(.1, .1), pkg.Normal(10, 5), pkg.Uniform(0, 1)]
    om = omodels.one_timescale_and_osc_model(data, t, prior=po, summary_method="psd", freqlims=(0.5, hi * fs))
    gm = pkg.one_timescale_and_osc_model(data, t, "abc", prior=pp, summary_method="psd", freqlims=(0.5, hi * fs))
    theta = np.array([[0.05, 10.0, 0.9], [0.1, 12.0, 0.5]])
    want, ws = cport.abc_distances(om, theta, seed=1, generation=1, return_stats=True)
    got, gs = gm.gpu_model(ctx).distances(theta, seed=1, generation=1, return_stats=True)
    assert np.max(np.abs(gs - ws) / ws.max(axis=1, keepdims=True)) <= 1e-5
    assert np.all(np.abs(got - want) <= 1e-3 * want)
    x = np.random.default_rng(0).standard_normal((2, 500))
    assert np.all(np.isfinite(pkg.summary.comp_psd(x, fs, ctx=ctx)[0]))          # the context is still healthy
