"""The C restatement (oracle/c/itjl_oracle.c, used for CPU baselines and posterior parity)
must agree with the NumPy oracle on every model variant."""
import numpy as np
import pytest

from oracle import cport, models, ou

DT = 0.002


def _time(n):
    return DT * np.arange(1, n + 1)


def test_plain_acf_matches_numpy_oracle():
    n = 1500
    data = ou.generate_ou_process(0.15, 1.0, DT, _time(n)[-1], 2, seed=123, n_t=n)
    m = models.one_timescale_model(data, _time(n), prior=[models.Uniform(0.05, 5)])
    th = np.array([[0.15], [1.0], [0.05], [-0.5], [-0.001]])
    dc = cport.abc_distances(m, th, seed=5, generation=1)
    dn = np.array([models.generate_data_and_reduce(m, x, seed=5, generation=1, particle=i)
                   for i, x in enumerate(th)])
    assert np.isnan(dc[-1]) and np.isnan(dn[-1])            # diverging trajectory -> NaN
    assert np.allclose(dc[:-1], dn[:-1], rtol=1e-10, atol=1e-14)
    # injected noise / offsets / EM variant
    rng = np.random.default_rng(0)
    u0 = rng.standard_normal((3, 2)); nz = rng.standard_normal((3, 2, n))
    m.update = "em"
    dc = cport.abc_distances(m, th[:3], u0=u0, noise=nz)
    dn = [models.generate_data_and_reduce(m, th[i], u0=u0[i], noise=nz[i]) for i in range(3)]
    assert np.allclose(dc, dn, rtol=1e-10, atol=1e-14)
    m.update = "exact"
    d_off = cport.abc_distances(m, th[1:3], seed=5, generation=1, particle_offset=1)
    assert np.array_equal(d_off, cport.abc_distances(m, th[:3], seed=5, generation=1)[1:3])
    assert np.array_equal(cport.abc_distances(m, th[:4], seed=5, generation=1, n_threads=1),
                          cport.abc_distances(m, th[:4], seed=5, generation=1, n_threads=3))


def test_missing_acf_matches_numpy_oracle():
    n = 1200
    data = ou.generate_ou_process(0.2, 1.0, DT, _time(n)[-1], 3, seed=7, n_t=n)
    for r in range(3):
        data[r, 100 * r + 50:100 * r + 170] = np.nan
    m = models.one_timescale_with_missing_model(data, _time(n), prior=[models.Uniform(0.05, 5)])
    th = np.array([[0.2], [0.7], [0.06]])
    dc = cport.abc_distances(m, th, seed=9, generation=2)
    dn = [models.generate_data_and_reduce(m, x, seed=9, generation=2, particle=i) for i, x in enumerate(th)]
    assert np.allclose(dc, dn, rtol=1e-10, atol=1e-14)


@pytest.mark.parametrize("summary_method", ["psd", "acf"])
def test_osc_matches_numpy_oracle(summary_method):
    n = 1500
    data = ou.generate_ou_with_oscillation([0.05, 10.0, 0.9], DT, _time(n)[-1], 3, 0.5, 2.0, seed=11, n_t=n)
    prior = [models.Normal(.1, .1), models.Normal(10, 5), models.Uniform(0, 1)]
    m = models.one_timescale_and_osc_model(data, _time(n), prior=prior, summary_method=summary_method)
    th = np.array([[0.05, 10, 0.9], [0.1, 12, 0.5], [0.02, 8, 1.2], [0.03, 9.5, -0.2]])
    dc = cport.abc_distances(m, th, seed=5, generation=3)
    dn = [models.generate_data_and_reduce(m, x, seed=5, generation=3, particle=i) for i, x in enumerate(th)]
    assert np.allclose(dc, dn, rtol=1e-9, atol=1e-14)


def test_statistics_output_matches_numpy_oracle():
    """itjl_oracle_abc_stats: the per-particle trial-mean statistics behind each distance (used by the
    many-particle GPU parity test of the tensor-core kernel) equal the NumPy oracle's."""
    n = 1500
    data = ou.generate_ou_process(0.15, 1.0, DT, _time(n)[-1], 3, seed=123, n_t=n)
    data_m = data.copy(); data_m[1, 200:320] = np.nan
    for m in (models.one_timescale_model(data, _time(n), prior=[models.Uniform(0.05, 5)]),
              models.one_timescale_with_missing_model(data_m, _time(n), prior=[models.Uniform(0.05, 5)])):
        th = np.array([[0.15], [1.0], [0.05], [-0.001]])
        d, s = cport.abc_distances(m, th, seed=8, generation=2, return_stats=True)
        assert s.shape == (4, m.n_lags) and np.isnan(s[-1]).all() and np.isnan(d[-1])
        for i in range(3):
            ref = models.summary_stats(m, models.generate_data(m, th[i], seed=8, generation=2, particle=i))
            assert np.allclose(s[i], ref, rtol=0, atol=1e-12)
            assert np.isclose(d[i], np.mean((ref - m.data_sum_stats) ** 2), rtol=1e-10)
        assert np.array_equal(d, cport.abc_distances(m, th, seed=8, generation=2), equal_nan=True)


def test_osc_with_missing_matches_numpy_oracle_and_uncovered_models_are_refused():
    n = 1600
    data = ou.generate_ou_with_oscillation([0.05, 10.0, 0.9], DT, _time(n)[-1], 4, 0.5, 2.0, seed=17, n_t=n)
    for r in range(4):
        data[r, 90 * r + 40:90 * r + 200] = np.nan
    prior = [models.Normal(.1, .1), models.Normal(10, 5), models.Uniform(0, 1)]
    m = models.one_timescale_and_osc_with_missing_model(data, _time(n), prior=prior)
    th = np.array([[0.05, 10, 0.9], [0.1, 12, 0.5], [0.02, 8, 1.2]])
    dc, sc = cport.abc_distances(m, th, seed=4, generation=2, return_stats=True)
    for i, x in enumerate(th):
        ref = models.summary_stats(m, models.generate_data(m, x, seed=4, generation=2, particle=i))
        assert np.allclose(sc[i], ref, rtol=0, atol=1e-12)
        assert np.isclose(dc[i], models.generate_data_and_reduce(m, x, seed=4, generation=2, particle=i), rtol=1e-9)
    plain = models.one_timescale_model(np.nan_to_num(data), _time(n), prior=[models.Uniform(0.05, 5)])
    with pytest.raises(ValueError):
        cport.abc_distances(models.with_combined_distance(plain), np.array([[0.3]]))
