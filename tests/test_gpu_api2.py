"""GPU tests of the round-2 additions to the C ABI: per-context tuning (no environment switches), the context
lock, strided / device-resident inputs of the summary and acw calls, and the multi-GPU context on one device."""
import threading

import numpy as np
import pytest

from oracle import acw as oacw
from oracle import ou as oou
from oracle import summary as osum

pytestmark = pytest.mark.gpu


def _ou(n_trials, n_t, seed=1, tau=0.05, dt=0.002):
    return oou.generate_ou_process(tau, 1.0, dt, n_t * dt, n_trials, seed=seed, n_t=n_t)


def test_tuning_is_per_context_and_read_at_model_creation(pkg, ctx):
    data = _ou(3, 2000)
    t = 0.002 * np.arange(1, 2001)
    mk = lambda: pkg.one_timescale_model(data, t, "abc", prior=[pkg.Uniform(0.01, 1)])
    other = pkg.Context(0)
    try:
        ctx.reset_tuning()
        other.set_tuning(abc_tc=0)
        assert mk().gpu_model(ctx).kernel_info()[0] == "k_abc_acf_tc"
        assert mk().gpu_model(other).kernel_info()[0] == "k_abc_acf"           # the other context's knob
        assert ctx.get_tuning().abc_tc == -1
        with ctx.tuned(abc_tc=0):
            g = mk().gpu_model(ctx)
        assert g.kernel_info()[0] == "k_abc_acf" and ctx.get_tuning().abc_tc == -1
        with pytest.raises(pkg.ItjlError):
            ctx.set_tuning(tc_terms=7)
        with pytest.raises(KeyError):
            ctx.set_tuning(no_such_knob=1)
    finally:
        other.close()


def test_environment_is_ignored(pkg, ctx, monkeypatch):
    """Round 1 switched kernels on ITJL_* environment variables; the library no longer reads any."""
    data = _ou(3, 2000)
    t = 0.002 * np.arange(1, 2001)
    monkeypatch.setenv("ITJL_ABC_TC", "0")
    monkeypatch.setenv("ITJL_ABC_DEBUG", "3")
    monkeypatch.setenv("ITJL_TC_TERMS", "1")
    g = pkg.one_timescale_model(data, t, "abc", prior=[pkg.Uniform(0.01, 1)]).gpu_model(ctx)
    assert g.kernel_info()[0] == "k_abc_acf_tc"
    theta = np.array([[0.05], [0.3]])
    monkeypatch.delenv("ITJL_ABC_TC"); monkeypatch.delenv("ITJL_ABC_DEBUG"); monkeypatch.delenv("ITJL_TC_TERMS")
    g2 = pkg.one_timescale_model(data, t, "abc", prior=[pkg.Uniform(0.01, 1)]).gpu_model(ctx)
    assert np.array_equal(g.distances(theta, seed=1), g2.distances(theta, seed=1))


def test_context_lock_serialises_concurrent_callers(pkg, ctx):
    """callable from any host thread, one in-flight call per context (guarded by the library)."""
    data = _ou(4, 3000)
    t = 0.002 * np.arange(1, 3001)
    g = pkg.one_timescale_model(data, t, "abc", prior=[pkg.Uniform(0.01, 1)]).gpu_model(ctx)
    theta = np.random.default_rng(0).uniform(0.02, 0.9, size=(400, 1))
    want = g.distances(theta, seed=4, generation=2)
    acf_want = pkg.comp_ac_fft(data, 100, ctx=ctx)
    out, errs = {}, []

    def worker(i):
        try:
            for _ in range(5):
                if i % 2:
                    out[i] = g.distances(theta, seed=4, generation=2)
                else:
                    out[i] = pkg.comp_ac_fft(data, 100, ctx=ctx)
        except Exception as e:                                   # noqa: BLE001
            errs.append(e)
    th = [threading.Thread(target=worker, args=(i,)) for i in range(6)]
    [x.start() for x in th]; [x.join() for x in th]
    assert not errs
    for i in range(6):
        assert np.array_equal(out[i], want if i % 2 else acf_want)


@pytest.mark.parametrize("missing", [False, True])
def test_strided_inputs_need_no_host_transpose(pkg, ctx, missing):
    """slice_stride / time_stride of itjl_acf_strided / itjl_acw_strided: a C-ordered (trials, time)
    NumPy matrix is the memory layout of a Julia matrix with time down the columns (dims = 1)."""
    x = _ou(37, 1234, seed=3)
    if missing:
        x[3, 100:300] = np.nan; x[20, ::7] = np.nan
    ref = (osum.comp_ac_time_missing if missing else osum.comp_ac_fft)(x, 80)
    fn = pkg.comp_ac_time_missing if missing else pkg.comp_ac_fft
    xc, xf = np.ascontiguousarray(x), np.asfortranarray(x)
    a, b = fn(xc, 80, ctx=ctx), fn(xf, 80, ctx=ctx)
    assert np.array_equal(a, b) and np.max(np.abs(a - ref)) < 1e-12
    types = ["acw0", "acw50", "acweuler", "tau"]
    r1 = pkg.acw(xc, 500.0, acwtypes=types, ctx=ctx)
    r2 = pkg.acw(xf, 500.0, acwtypes=types, ctx=ctx)
    r3 = pkg.acw(np.ascontiguousarray(x.T), 500.0, acwtypes=types, dims=1, ctx=ctx)      # (time, trials), dims = 1
    want = oacw.acw(x, 500.0, acwtypes=types)
    for r in (r1, r2, r3):
        assert r.n_lags == want["n_lags"]
        for k in range(len(types)):
            assert np.array_equal(r.acw_results[k], r1.acw_results[k], equal_nan=True)
    assert np.array_equal(r1.acw_results[0], np.where(want["k0"] >= 0, want["k0"] * (1.0 / 500.0), np.nan), equal_nan=True)
    assert np.array_equal(r3.acf.T, r1.acf)


def test_device_resident_data_for_repeated_acw_calls(pkg, ctx):
    x = _ou(64, 4000, seed=9, tau=0.1)
    dd = pkg.upload(x, ctx=ctx)
    launches0 = ctx.launch_count
    a = pkg.acw(dd, 500.0, acwtypes=["acw0", "acw50"])
    b = pkg.acw(dd, 500.0, acwtypes=["tau", "auc"], n_lags=300)
    c = pkg.acw(dd, 500.0, acwtypes=["acw0"], average_over_trials=True)
    assert ctx.launch_count > launches0
    for got, kw in ((a, dict(acwtypes=["acw0", "acw50"])), (b, dict(acwtypes=["tau", "auc"], n_lags=300)),
                    (c, dict(acwtypes=["acw0"], average_over_trials=True))):
        want = pkg.acw(x, 500.0, ctx=ctx, **kw)
        assert got.n_lags == want.n_lags and np.array_equal(got.acf, want.acf)
        gr = got.acw_results if isinstance(got.acw_results, list) else [got.acw_results]
        wr = want.acw_results if isinstance(want.acw_results, list) else [want.acw_results]
        for u, v in zip(gr, wr):
            assert np.array_equal(u, v, equal_nan=True)
    # another data call on the context replaces the resident matrix: the library says so instead of using stale data
    dd2 = pkg.upload(x[:8], ctx=ctx)
    pkg.comp_ac_fft(x[:2], 10, ctx=ctx)
    with pytest.raises(pkg.ItjlError, match="no resident data"):
        pkg.acw(dd2, 500.0, acwtypes=["acw0"])


def test_multi_gpu_context_on_one_device_gives_the_single_gpu_results(pkg, ctx):
    """itjl_ctx_create_multi with one device exercises the whole sharded path (sub-context, sub-model, gather
    buffer, compaction) without NCCL; tests with 2+ GPUs are benchmarks/multi_gpu_check.py (gpurun --gpus 2)."""
    data = _ou(5, 3000, seed=2, tau=0.2)
    t = 0.002 * np.arange(1, 3001)
    model = pkg.one_timescale_model(data, t, "abc", prior=[pkg.Uniform(0.01, 2)])
    theta = np.random.default_rng(1).uniform(0.02, 1.9, size=(333, 1))
    g = model.gpu_model(ctx)
    want = g.generation(theta, 0.05, 20, seed=6, generation=3, particle_offset=11)
    multi = pkg.Context(devices=[0])
    try:
        assert multi.world() == (0, 1, 1)
        gm = model.gpu_model(multi)
        got = gm.generation(theta, 0.05, 20, seed=6, generation=3, particle_offset=11)
        assert np.array_equal(got[0], want[0]) and np.array_equal(got[1][:got[2]], want[1][:want[2]]) and got[2:] == want[2:]
        assert np.array_equal(gm.distances(theta, seed=6, generation=3, particle_offset=11), want[0])
        d, stats = gm.distances(theta[:5], seed=6, generation=3, particle_offset=11, return_stats=True)   # first device
        assert np.array_equal(d, want[0][:5]) and stats.shape == (5, model.n_lags)
        assert gm.kernel_info() == g.kernel_info() and gm.work() == g.work()
        rng = np.random.default_rng(3)
        th_prev = rng.uniform(0.1, 1, (50, 1)); th = rng.uniform(0.1, 1, (70, 1)); w = np.full(50, 0.02)
        cov = np.array([[0.1]])
        assert np.array_equal(pkg.calc_weights(th_prev, th, cov, w, [pkg.Uniform(0, 2)], ctx=multi),
                              pkg.calc_weights(th_prev, th, cov, w, [pkg.Uniform(0, 2)], ctx=ctx))
        assert np.max(np.abs(pkg.comp_ac_fft(data, 50, ctx=multi) - pkg.comp_ac_fft(data, 50, ctx=ctx))) == 0
    finally:
        multi.close()
    with pytest.raises(pkg.ItjlError):
        pkg.Context(devices=[0, 0])
