"""The oracle's PSD-side reducers (oracle/knee.py) against the reference's own tests for them
(/root/reference/test/test_utils.jl:8-124) and against first principles.  CPU only."""
import numpy as np

from oracle import knee, ou as oou, summary as osum


def test_ideal_lorentzian_knee():
    """test_utils.jl:11-19: rtol 1e-3."""
    freqs = np.arange(1, 1001) * 0.1 / 1000.0
    psd = 100.0 / (1 + (freqs / 0.015) ** 2)
    amp, k = knee.find_knee_frequency(psd, freqs)
    assert abs(k - 0.015) < 1e-3 * 0.015 and abs(amp - 100.0) < 1e-6


def test_noisy_lorentzian_and_frequency_ranges():
    """test_utils.jl:22-33 (rtol 0.05) and :53-66 (rtol 0.01)."""
    rng = np.random.default_rng(123)
    freqs = np.arange(0, 1001) * 0.1 / 1000.0
    base = 100.0 / (1 + (freqs / 0.010) ** 2)
    k = knee.find_knee_frequency(base + rng.standard_normal(freqs.size) * 5.0, freqs)[1]
    assert abs(k - 0.010) < 0.05 * 0.010
    k1 = knee.find_knee_frequency(base, freqs, min_freq=0.1 / 1000.0, max_freq=50.0)[1]
    k2 = knee.find_knee_frequency(base, freqs, min_freq=1.0 / 1000.0, max_freq=20.0 / 1000.0)[1]
    assert abs(k1 - k2) < 0.01 * k1 and abs(k1 - 0.010) < 0.01 * 0.010


def test_edge_cases_return_the_band_edge():
    """test_utils.jl:69-82: flat and white spectra give freqs[end]."""
    freqs = np.arange(0, 1001) * 0.1 / 1000.0
    assert np.isclose(knee.find_knee_frequency(np.ones(freqs.size), freqs)[1], freqs[-1])
    # white noise: the reference asserts rtol 0.01 on ONE realisation (Random.seed!(666)); the objective is flat in the
    # knee there, so over several realisations the argmin stays within a few percent of the band edge
    for seed in range(8):
        k = knee.find_knee_frequency(np.random.default_rng(seed).standard_normal(freqs.size), freqs)[1]
        assert abs(k - freqs[-1]) < 0.06 * freqs[-1]


def test_ou_process_knee_gives_the_timescale():
    """test_utils.jl:85-98: tau = 10, dt = 1, 100 trials of 5000 samples, rtol 0.1."""
    x = oou.generate_ou_process(10.0, 1.0, 1.0, 5000.0, 100, seed=4, n_t=5000)
    psd, freqs = osum.comp_psd_periodogram(x, 1.0)
    k = knee.find_knee_frequency(psd.mean(axis=0), freqs, min_freq=freqs[0], max_freq=freqs[-1])[1]
    assert abs(knee.tau_from_knee(k) - 10.0) < 1.0


def test_oscillation_peak_and_gaussian_fit():
    f = np.arange(1, 400) * 0.25
    lor = 10.0 / (1 + (f / 8.0) ** 2)
    p = lor + 3.0 * np.exp(-(f - 30.0) ** 2 / (2 * 2.0 ** 2))
    u = knee.find_knee_frequency(p, f)
    pk = knee.find_oscillation_peak(p - knee.lorentzian(f, u), f, f[0], f[-1])
    assert abs(pk - 30.0) <= 0.5
    g = knee.fit_gaussian(p - lor, f, 30.0)
    assert abs(g[0] - 3.0) < 1e-6 and abs(g[1] - 30.0) < 1e-6 and abs(g[2] - 2.0) < 1e-6
    # a monotone spectrum has no strict local maximum
    assert np.isnan(knee.find_oscillation_peak(lor, f, f[0], f[-1]))
    # prominence is measured against the larger of the two one-sided global minima (utils.jl:741-746)
    s = np.array([5.0, 1.0, 4.0, 3.5, 6.0, 0.0, 2.0, 1.5])
    assert knee.find_oscillation_peak(s, np.arange(8.0), 0.0, 7.0) == 4.0
    k, peaks = knee.fooof_fit(p, f, max_peaks=1, return_only_knee=False)
    assert abs(k - 8.0) < 0.05 * 8.0 and abs(peaks[0][0] - 30.0) < 0.5
