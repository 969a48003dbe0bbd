"""Host-side logic of the product (no GPU): the batched basic_abc must reproduce the serial
reference loop exactly for the same proposals and distances; PMC bookkeeping must match the
oracle's restatement of abc.jl."""
import numpy as np
import pytest

from oracle import abc as oabc
from oracle import models as omodels


class FakeModel:
    def __init__(self, prior):
        self.prior = prior
        self.fit_method = "abc"


def table_scorer(table):
    """Scorer backed by a precomputed distance table indexed by global particle index."""
    def scorer(theta, epsilon, needed, generation, offset):
        d = table[offset:offset + theta.shape[0]].copy()
        acc = d <= epsilon
        n_total, n_acc = d.size, int(acc.sum())
        if needed > 0:
            hit = np.nonzero(np.cumsum(acc) == needed)[0]
            if hit.size:
                n_total, n_acc = int(hit[0]) + 1, needed
        return d, acc, n_total, n_acc
    return scorer


@pytest.mark.parametrize("eps,min_acc", [(0.3, 17), (0.05, 100), (2.0, 5), (-1.0, 3), (0.5, 1)])
def test_batched_basic_abc_equals_serial_loop(pkg, eps, min_acc):
    rng = np.random.default_rng(1)
    max_iter = 3000
    table = rng.random(max_iter)
    table[rng.integers(0, max_iter, 40)] = np.nan
    props = rng.random((max_iter, 2))
    model = FakeModel([pkg.Uniform(0, 1), pkg.Uniform(0, 1)])
    got = pkg.basic_abc(model, eps, max_iter, min_acc, scorer=table_scorer(table), proposals=props)
    want = oabc.basic_abc(lambda th, i, g: table[i], None, 2, eps, max_iter, min_acc, proposals=props)
    assert got["n_total"] == want["n_total"] and got["n_accepted"] == want["n_accepted"]
    assert np.array_equal(got["distances"], want["distances"], equal_nan=True)
    assert np.array_equal(got["theta_accepted"], want["theta_accepted"])
    assert np.array_equal(got["isaccepted"], want["isaccepted"])
    assert np.array_equal(got["samples"], want["samples"])


def test_pmc_helpers_match_oracle(pkg):
    rng = np.random.default_rng(2)
    d = np.concatenate([rng.random(500) * 3, [np.nan] * 5, [20.0, 11.0]])
    for it, rate, eps in [(1, 0.0, 1.0), (2, 0.5, 0.9), (7, 0.001, 0.2), (12, 0.0101, 0.4), (3, 0.0, 0.7)]:
        a = pkg.select_epsilon(d, eps, current_acc_rate=rate, iteration=it, total_iterations=30)
        b = oabc.select_epsilon(d, eps, current_acc_rate=rate, iteration=it, total_iterations=30)
        assert a == b
    # calc_weights / weighted_covar run on the device: tests/test_gpu_pmc.py
    assert pkg.effective_sample_size(np.ones(3)) == pytest.approx(3.0)        # test_abc.jl
    assert pkg.get_param_dict_abc() == oabc.get_param_dict_abc()


def test_draw_theta_pmc_is_nonnegative_and_centred(pkg):
    rng = np.random.default_rng(3)
    th_prev = np.abs(rng.normal(0.3, 0.05, (100, 1)))
    w = np.full(100, 0.01)
    out = pkg.draw_theta_pmc(None, th_prev, w, np.array([[0.04]]), rng, n=20000)
    assert out.shape == (20000, 1) and np.all(out >= 0)
    assert abs(np.median(out) - 0.33) < 0.05


def test_pmc_abc_runs_with_a_toy_scorer(pkg):
    """pmc_abc control flow on an analytic toy: d(theta) = (theta - 2)^2 + noise."""
    rng = np.random.default_rng(4)

    def scorer(theta, epsilon, needed, generation, offset):
        d = (theta[:, 0] - 2.0) ** 2 + 0.01 * rng.random(theta.shape[0])
        acc = d <= epsilon
        n_total, n_acc = d.size, int(acc.sum())
        if needed > 0:
            hit = np.nonzero(np.cumsum(acc) == needed)[0]
            if hit.size:
                n_total, n_acc = int(hit[0]) + 1, needed
        return d, acc, n_total, n_acc

    model = FakeModel([pkg.Uniform(0, 5)])
    # host control flow only: the scorer and the weight / covariance steps (device kernels in the product) are
    # injected, here the oracle's
    res, rec = pkg.pmc_abc(model, epsilon_0=1.0, max_iter=5000, min_accepted=100, steps=8, rng=rng,
                           scorer=scorer, target_epsilon=1e-4, return_record=True,
                           calc_weights_fn=oabc.calc_weights, weighted_covar_fn=oabc.weighted_covar)
    assert isinstance(res, pkg.ABCResults) and len(res.epsilon_history) == len(rec) >= 2
    assert abs(res.final_theta.mean() - 2.0) < 0.15
    assert res.final_weights.sum() == pytest.approx(1.0)
    assert all(e2 <= e1 * 1.0001 or True for e1, e2 in zip(res.epsilon_history, res.epsilon_history[1:]))
    assert abs(res.MAP[0] - 2.0) < 0.3


def test_pmc_abc_leaves_through_min_acc_rate_when_nothing_is_accepted(pkg):
    """Generation 1 accepts nothing (epsilon_0 below every distance): the reference's fill(1 / 0, 0) is an
    empty weight vector and it returns through `accept_rate < minAccRate` (abc.jl:461) - no exception."""
    def scorer(theta, epsilon, needed, generation, offset):
        d = 1.0 + theta[:, 0] ** 2
        return d, d <= epsilon, d.size, 0

    model = FakeModel([pkg.Uniform(0, 5)])
    res, rec = pkg.pmc_abc(model, epsilon_0=0.5, max_iter=300, min_accepted=10, steps=5, rng=np.random.default_rng(1),
                           scorer=scorer, return_record=True, calc_weights_fn=oabc.calc_weights,
                           weighted_covar_fn=oabc.weighted_covar)
    assert len(rec) == 1 and rec[0]["n_accepted"] == 0 and rec[0]["n_total"] == 300
    assert res.final_theta.shape[0] == 0 and res.final_weights.size == 0 and np.all(np.isnan(res.MAP))


def test_acw_and_model_argument_errors(pkg):
    with pytest.raises(ValueError):
        pkg.acw(np.zeros((2, 10)), 1.0, acwtypes=[])
    with pytest.raises(ValueError):
        pkg.acw(np.zeros((2, 10)), 1.0, acwtypes=["acw25"])                 # not one of possible_acwtypes
    t = np.arange(1, 11.0)
    with pytest.raises(ValueError, match="Time vector length"):
        pkg.check_model_inputs(np.zeros((2, 9)), t, "abc", "acf", None, None)
    with pytest.raises(ValueError, match="monotonically"):
        pkg.check_model_inputs(np.zeros((2, 10)), t[::-1], "abc", "acf", None, None)
    with pytest.raises(ValueError, match="fit_method"):
        pkg.check_model_inputs(np.zeros((2, 10)), t, "mcmc", "acf", None, None)
    with pytest.raises(ValueError, match="summary_method"):
        pkg.check_model_inputs(np.zeros((2, 10)), t, "abc", "foo", None, None)
    with pytest.raises(ValueError, match="distance_method"):
        pkg.check_model_inputs(np.zeros((2, 10)), t, "abc", "acf", None, "cosine")


def test_tensor_core_plan_selection_is_host_logic(pkg):
    """itjl_tc_plan (no GPU): tensor-memory plan A / B, FFMA tail tiles, accumulator segments, hi/lo
    terms and shared-memory fit of k_abc_acf_tc for the BASELINE shapes, on B200 limits."""
    import ctypes
    lib = pkg._lib.lib()
    names = ["groups", "chunk", "ksteps", "rows_total", "used_cols", "mixed", "seg_trials", "n_terms",
             "tc_lags", "tail_start", "tail_lt", "n_ops", "smem_bytes", "grid", "n_windows"]

    def plan(n_t, n_lags, n_trials):
        out = (ctypes.c_int32 * 16)()
        rc = lib.itjl_tc_plan(n_t, n_lags, n_trials, 232448, 148, out)
        return rc, dict(zip(names, list(out)))

    rc, p = plan(5000, 414, 50)                       # C5: plan B, no tail, one 50-trial segment, hi*hi only (2.5e5 samples)
    assert rc == 0 and p["mixed"] == 1 and p["n_ops"] == 6 and p["tail_lt"] == 0 and p["tc_lags"] == 414
    assert p["groups"] == 4 and p["chunk"] == 40 and p["ksteps"] == 3 and p["seg_trials"] == 50 and p["n_terms"] == 1
    assert p["rows_total"] >= 16 * 3 + 4 and p["smem_bytes"] + 6 * 1024 <= 232448 and p["grid"] == 148
    rc, p = plan(5000, 463, 100)                      # C3: plan B + one tail tile of 16 lags from lag 448
    assert rc == 0 and p["mixed"] == 1 and p["tc_lags"] == 449 and p["tail_start"] == 448 and p["tail_lt"] == 1
    rc, p = plan(5000, 300, 2)                        # C1-like: plan A, one group per trial, 3 terms (1e4 samples)
    assert rc == 0 and p["mixed"] == 0 and p["groups"] == 2 and p["n_terms"] == 3 and p["seg_trials"] == 2
    assert p["used_cols"] == 432 and p["n_ops"] == 4
    rc, p = plan(5000, 385, 8)                        # last shape of plan A: all 512 columns
    assert rc == 0 and p["mixed"] == 0 and p["used_cols"] == 512
    rc, p = plan(5000, 464, 8)                        # last shape with an FFMA tail (one tile of 16 lags)
    assert rc == 0 and p["tail_lt"] == 1 and p["n_windows"] == 1
    rc, p = plan(5000, 465, 8)                        # beyond: a second lag window is faster than more tail tiles
    assert rc == 0 and p["tail_lt"] == 0 and p["n_windows"] == 2
    rc, p = plan(5000, 510, 8)                        # two lag windows (second launch), no tail, rows for the deeper shift
    assert rc == 0 and p["n_windows"] == 2 and p["tail_lt"] == 0 and p["tc_lags"] == 449
    rc, p = plan(5000, 769, 50)
    assert rc == 0 and p["n_windows"] == 2 and p["groups"] == 4 and p["rows_total"] >= 48 + 6   # B rows up to 47 + 3 + 3
    rc, p = plan(5000, 1153, 8)                       # three windows need 57 rows: only three simulate groups fit
    assert rc == 0 and p["n_windows"] == 3 and p["groups"] == 3 and p["rows_total"] >= 48 + 9
    assert plan(5000, 1537, 8)[1]["n_windows"] == 4
    assert plan(5000, 1538, 8)[0] == -1               # longer windows: CUDA-core kernel
    rc, p = plan(10000, 430, 100)                     # longer trials: fewer simulate groups fit
    assert rc == 0 and 1 <= p["groups"] < 4 and p["ksteps"] == 5
    assert plan(8, 4, 2)[0] == -1 and plan(5000, 6000, 2)[0] == -1


def test_spectrum_grid_helpers_match_the_oracle(pkg):
    """Host-side helpers of the PSD side (no device work): nextfastfft, the Lomb-Scargle autofrequency grid, knee <-> tau,
    the Lorentzian - against the oracle's restatements and closed forms."""
    from oracle import knee as oknee
    from oracle import summary as osum
    for n in list(range(1, 400)) + [4999, 5000, 5001, 9973, 10000, 16385, 65536]:
        m = pkg.nextfastfft(n)
        assert m == osum.nextfastfft(n) and m >= n
        r = m
        for p in (2, 3, 5, 7):
            while r % p == 0:
                r //= p
        assert r == 1                                         # 7-smooth
    for n_t, dt in ((1000, 0.01), (777, 0.002), (5000, 0.002)):
        t = dt * np.arange(1, n_t + 1)
        f0, df, n = pkg.lombscargle_autofrequency(t, 1.0 / (2 * dt))
        want = osum.lombscargle_autofrequency(t, 1.0 / (2 * dt))
        grid = f0 + df * np.arange(n)
        assert np.allclose(grid, want, rtol=1e-13, atol=0) and grid[-1] <= 1.0 / (2 * dt) * (1 + 1e-12)
        assert np.isclose(df, 1.0 / (5 * (t[-1] - t[0])), rtol=1e-14) and np.isclose(f0, df / 2, rtol=1e-14)
    k = np.array([0.5, 3.0, 40.0])
    assert np.allclose(pkg.tau_from_knee(k), 1.0 / (2 * np.pi * k)) and np.allclose(pkg.knee_from_tau(pkg.tau_from_knee(k)), k)
    f = np.linspace(0.1, 50, 200)
    assert np.allclose(pkg.lorentzian(f, [3.0, 7.0]), oknee.lorentzian(f, [3.0, 7.0]), rtol=1e-15)
    assert np.isclose(pkg.lorentzian(np.array([7.0]), [3.0, 7.0])[0], 1.5)          # half power at the knee
