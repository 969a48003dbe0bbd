"""Oracle: summary statistics (ACF / NaN-aware ACF / Hamming periodogram).

Follows /root/reference/src/stats/summary.jl:
  comp_ac_fft           :46-63 (vector), :79-89 (array)
  comp_psd_adfriendly   :183-235
  comp_ac_time_missing  :455-484 (vector), :486-496 (array)
  comp_psd (periodogram) :111-160 (DSP.jl semantics restated, see comp_psd_periodogram_vec)
and /root/reference/src/stats/bat_autocor.jl:21-24 (_ac_next_pow_two).
Arrays are (trials, time): the reference's default dims=ndims(data).
"""
import numpy as np


def nextpow2(n):
    """nextpow(2, n)."""
    p = 1
    while p < n:
        p <<= 1
    return p


def _ac_next_pow_two(i):
    """bat_autocor.jl:21-24: smallest power of two >= i (i > 0)."""
    return nextpow2(i) if i > 0 else 1


def comp_ac_fft_vec(x, n_lags=None):
    """summary.jl:46-63."""
    x = np.asarray(x, dtype=np.float64)
    n = x.shape[0]
    if n_lags is None:
        n_lags = n
    xc = x - x.mean()
    n_pad = nextpow2(2 * n - 1)
    f = np.fft.fft(xc, n_pad)
    acov = np.real(np.fft.ifft(np.abs(f) ** 2))[:n] / n
    with np.errstate(invalid="ignore", divide="ignore"):
        ac = acov / acov[0]
    return ac[:n_lags]


def comp_ac_fft(data, n_lags=None):
    """summary.jl:79-89 on (trials, time) -> (trials, n_lags)."""
    data = np.atleast_2d(np.asarray(data, dtype=np.float64))
    return np.stack([comp_ac_fft_vec(r, n_lags) for r in data])


def comp_ac_direct_vec(x, n_lags):
    """Time-domain restatement of comp_ac_fft (identical up to fp64 rounding; the
    reference's own cross-check is test/test_summary_stats.jl:57-74)."""
    x = np.asarray(x, dtype=np.float64)
    n = x.shape[0]
    xc = x - x.mean()
    ac = np.array([np.dot(xc[:n - k], xc[k:]) for k in range(n_lags)])
    return ac / ac[0]


def comp_ac_time_missing_vec(x, n_lags=None):
    """summary.jl:455-484.  Note the reference's quirk: after NaNs are zeroed the
    valid-sample mean is subtracted from EVERY entry, so masked samples become
    -mean rather than 0 (:466-471)."""
    x = np.asarray(x, dtype=np.float64)
    n = x.shape[0]
    if n_lags is None:
        n_lags = n
    valid = ~np.isnan(x)
    z = np.where(valid, x, 0.0)
    with np.errstate(invalid="ignore", divide="ignore"):
        xm = z.sum() / valid.sum()
        z = z - xm
        ss = np.sum(z * z) / n
        ac = np.array([np.dot(z[:n - k], z[k:]) / n for k in range(n_lags)])
        return ac / ss


def comp_ac_time_missing(data, n_lags=None):
    """summary.jl:486-496."""
    data = np.atleast_2d(np.asarray(data, dtype=np.float64))
    return np.stack([comp_ac_time_missing_vec(r, n_lags) for r in data])


def hamming_window(n):
    """summary.jl:214."""
    return 0.54 - 0.46 * np.cos(2.0 * np.pi * np.arange(n) / (n - 1))


def comp_psd_adfriendly_vec(x, fs, demean=True):
    """summary.jl:204-235 -> (psd[2:n2/2], freqs[2:n2/2])."""
    x = np.asarray(x, dtype=np.float64)
    n = x.shape[0]
    n2 = 2 * _ac_next_pow_two(n)
    xd = x - x.mean() if demean else x
    w = hamming_window(n)
    xp = np.zeros(n2)
    xp[:n] = xd * w
    scale = 1.0 / (fs * np.sum(w ** 2))
    xf = np.fft.fft(xp)
    psd = np.abs(xf[:n2 // 2]) ** 2 * scale
    # AbstractFFTs.fftfreq(n2, fs)[1 : n2 / 2] = (0 : n2 / 2 - 1) * (fs / n2)
    freqs = np.arange(n2 // 2, dtype=np.float64) * (fs / n2)
    return psd[1:], freqs[1:]


def comp_psd_adfriendly(data, fs):
    """summary.jl:183-202 on (trials, time) -> ((trials, n2/2-1), freqs)."""
    data = np.atleast_2d(np.asarray(data, dtype=np.float64))
    rows = [comp_psd_adfriendly_vec(r, fs) for r in data]
    return np.stack([r[0] for r in rows]), rows[0][1]


# ---- comp_psd, method = "periodogram"  ---------------------------------------------
def nextfastfft(n):
    """DSP.jl `nextfastfft(n)` = nextprod([2, 3, 5, 7], n): the smallest 7-smooth integer >= n."""
    m = int(n)
    while True:
        r = m
        for p in (2, 3, 5, 7):
            while r % p == 0:
                r //= p
        if r == 1:
            return m
        m += 1


def comp_psd_periodogram_vec(x, fs):
    """summary.jl:146-155: `dsp.periodogram(x; fs=fs, window=dsp.hamming)` without the DC bin.
    DSP.jl 0.8 (third-party, not under /root/reference; restated from its documented behaviour, UNPINNED
    until tests/golden/reference_v1.json exists): no detrending; the window function is called with
    the signal length (symmetric Hamming, 0.54 - 0.46 cos(2 pi k / (n - 1))); nfft = nextfastfft(n), zero
    padded; one-sided real FFT; power = |X|^2 / (fs * sum(w^2)), doubled except at DC and - for even nfft -
    at Nyquist; freq = rfftfreq(nfft, fs)."""
    x = np.asarray(x, dtype=np.float64)
    n = x.shape[0]
    nfft = nextfastfft(n)
    w = hamming_window(n)
    xf = np.fft.rfft(x * w, nfft)
    p = np.abs(xf) ** 2 / (fs * np.sum(w ** 2))
    if nfft % 2 == 0:
        p[1:-1] *= 2.0
    else:
        p[1:] *= 2.0
    # AbstractFFTs.rfftfreq(nfft, fs) = (0 : nfft / 2) * (fs / nfft): the multiplier is formed first
    freqs = np.arange(nfft // 2 + 1, dtype=np.float64) * (fs / nfft)
    return p[1:], freqs[1:]


def comp_psd_periodogram(data, fs):
    """summary.jl:111-144 on (trials, time) -> ((trials, nfft/2), freqs)."""
    data = np.atleast_2d(np.asarray(data, dtype=np.float64))
    rows = [comp_psd_periodogram_vec(r, fs) for r in data]
    return np.stack([r[0] for r in rows]), rows[0][1]


# ---- Lomb-Scargle ---------------------------------------------------------------------
def lombscargle_autofrequency(times, maximum_frequency, samples_per_peak=5):
    """LombScargle.jl 1.x `autofrequency(times; samples_per_peak = 5, maximum_frequency)`: df = 1 / (5 (max - min)),
    grid df/2 : df : maximum_frequency (third-party; restated from its documented behaviour, UNPINNED)."""
    times = np.asarray(times, dtype=np.float64)
    df = 1.0 / (samples_per_peak * (times.max() - times.min()))
    f0 = 0.5 * df
    n = int(np.floor((maximum_frequency - f0) / df + 1e-9)) + 1
    return f0 + df * np.arange(n)


def lombscargle_vec(times, signal, freqs):
    """summary.jl:254-259: `ls.lombscargle(ls.plan(times, signal; frequencies))` with LombScargle.jl's defaults
    normalization = :standard, fit_mean = true, center_data = true: the generalised (floating-mean) periodogram of
    Zechmeister & Kuerster (2009), eq. 4-15 with unit weights, evaluated exactly here.  For more than 200 samples on
    a range of frequencies the package itself uses the Press & Rybicki fast approximation of the same quantity
    (oversampling 5, Mfft 4) - its approximation error is not restated.  UNPINNED."""
    t = np.asarray(times, dtype=np.float64)
    y = np.asarray(signal, dtype=np.float64)
    y = y - y.mean()
    w = 1.0 / y.shape[0]
    YY = w * np.sum(y * y)
    out = np.empty(len(freqs))
    for lo in range(0, len(freqs), 256):
        f = np.asarray(freqs[lo:lo + 256], dtype=np.float64)
        ph = 2.0 * np.pi * np.outer(f, t)
        c, s_ = np.cos(ph), np.sin(ph)
        C, S = w * c.sum(axis=1), w * s_.sum(axis=1)
        YC, YS = w * (c @ y), w * (s_ @ y)
        CCh, CSh = w * np.sum(c * c, axis=1), w * np.sum(c * s_, axis=1)
        CC, SS, CS = CCh - C * C, (1.0 - CCh) - S * S, CSh - C * S
        D = CC * SS - CS * CS
        with np.errstate(invalid="ignore", divide="ignore"):
            pw = (SS * YC * YC + CC * YS * YS - 2.0 * CS * YC * YS) / (YY * D)
            # degenerate frequencies (exactly Nyquist on a regular grid: every sin(omega t) vanishes, D = 0 / 0): one basis
            # function is left - the power is the projection on it
            deg = ~(D > 1e-12 * (CC + SS) ** 2)
            one = np.where(SS < CC, YC * YC / (YY * CC), YS * YS / (YY * SS))
        out[lo:lo + 256] = np.where(deg, one, pw)
    return out


def comp_psd_lombscargle(times, data, nanmask, dt):
    """summary.jl:321-356 on (trials, time): common grid from the FULL time vector, every trial's valid samples."""
    data = np.atleast_2d(np.asarray(data, dtype=np.float64))
    nanmask = np.atleast_2d(np.asarray(nanmask, dtype=bool))
    times = np.asarray(times, dtype=np.float64)
    freqs = lombscargle_autofrequency(times, 1.0 / (2.0 * dt))
    rows = [lombscargle_vec(times[~m], r[~m], freqs) for r, m in zip(data, nanmask)]
    return np.stack(rows), freqs
