"""Oracle vs the committed golden fixtures (CPU), and the CUDA path vs the same fixtures (GPU)."""
import os

import numpy as np
import pytest

from oracle import acw, models, ou, philox, summary

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "oracle_v1.npz"))
DT, N_T = 0.002, 600
T = DT * np.arange(1, N_T + 1)


def test_oracle_reproduces_golden_vectors():
    got = np.array(philox.philox4x32(np.arange(4), 1, 2, philox.c3_word(0, 3), 7, 9)).astype(np.uint32)
    assert np.array_equal(got, G["philox_block"])
    assert np.allclose(philox.noise_normals(7, 3, 2, 1, 16), G["normals"], rtol=1e-14)
    x = ou.generate_ou_process(0.05, 1.0, DT, N_T * DT, 3, seed=7, n_t=N_T)
    assert np.allclose(x, G["ou"], rtol=1e-12, atol=1e-13)
    assert np.allclose(summary.comp_ac_fft(x, 50), G["acf"], atol=1e-13)
    assert np.allclose(summary.comp_ac_time_missing(G["data_missing"], 50), G["acf_missing"], atol=1e-13, equal_nan=True)
    p, f = summary.comp_psd_adfriendly(x, 1 / DT)
    assert np.allclose(p, G["psd"], rtol=1e-10) and np.allclose(f, G["psd_freqs"])
    m = models.one_timescale_model(x, T, prior=[models.Uniform(0.01, 1.0)])
    assert m.n_lags == int(G["model_n_lags"][0]) and np.allclose(m.data_sum_stats, G["model_stats"], atol=1e-13)
    d = [models.generate_data_and_reduce(m, th, seed=3, generation=1, particle=i) for i, th in enumerate(G["theta"])]
    assert np.allclose(d, G["dist"], rtol=1e-10)
    r = acw.acw(x, 1 / DT)
    assert np.array_equal(np.stack([r["k0"], r["k50"], r["ke"]]), G["acw_k"])
    assert np.allclose(r["auc"], G["acw_auc"], rtol=1e-12) and np.allclose(r["tau"], G["acw_tau"], rtol=1e-9)


@pytest.mark.gpu
def test_cuda_path_reproduces_golden_vectors(pkg, ctx):
    x = pkg.generate_ou_process(0.05, 1.0, DT, N_T * DT, 3, seed=7, ctx=ctx)
    assert np.max(np.abs(x - G["ou"])) <= 1e-5 * np.abs(G["ou"]).max()
    raw = pkg.generate_ou_process(0.05, 1.0, DT, N_T * DT, 2, standardize=False, seed=8, update="em", ctx=ctx)
    assert np.max(np.abs(raw - G["ou_em_raw"])) <= 1e-5 * np.abs(G["ou_em_raw"]).max()
    osc = pkg.generate_ou_with_oscillation([0.05, 10.0, 0.8], DT, N_T * DT, 2, 0.5, 2.0, seed=9, ctx=ctx)
    assert np.max(np.abs(osc - G["ou_osc"])) <= 2e-5 * np.abs(G["ou_osc"]).max()
    assert np.max(np.abs(pkg.comp_ac_fft(G["ou"], 50, ctx=ctx) - G["acf"])) < 1e-12
    assert np.max(np.abs(pkg.comp_ac_time_missing(G["data_missing"], 50, ctx=ctx) - G["acf_missing"])) < 1e-12
    p, f = pkg.comp_psd_adfriendly(G["ou"], 1 / DT, ctx=ctx)
    assert np.max(np.abs(p - G["psd"])) <= 1e-5 * G["psd"].max()
    m = pkg.one_timescale_model(G["ou"], T, "abc", prior=[pkg.Uniform(0.01, 1.0)])
    assert m.n_lags == int(G["model_n_lags"][0])
    d = m.gpu_model(ctx).distances(G["theta"], seed=3, generation=1)
    assert np.all(np.abs(d - G["dist"]) <= 2e-5 * np.sqrt(G["dist"]) + 1e-5 * G["dist"] + 1e-10)
    mm = pkg.one_timescale_with_missing_model(G["data_missing"], T, "abc", prior=[pkg.Uniform(0.01, 1.0)])
    dm = mm.gpu_model(ctx).distances(G["theta"], seed=3, generation=1)
    assert np.all(np.abs(dm - G["dist_missing"]) <= 2e-5 * np.sqrt(G["dist_missing"]) + 1e-5 * G["dist_missing"] + 1e-10)
    r = pkg.acw(G["ou"], 1 / DT, ctx=ctx)
    res = dict(zip(r.acwtypes, r.acw_results))
    assert np.array_equal(np.round(res["acw0"] / DT).astype(int), G["acw_k"][0])
    assert np.array_equal(np.round(res["acw50"] / DT).astype(int), G["acw_k"][1])
    assert np.array_equal(np.round(res["acweuler"] / DT).astype(int), G["acw_k"][2])
    assert np.allclose(res["auc"], G["acw_auc"], rtol=1e-6) and np.allclose(res["tau"], G["acw_tau"], rtol=1e-5)
    assert r.n_lags == int(G["acw_n_lags"][0])
