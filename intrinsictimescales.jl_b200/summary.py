"""Summary statistics on stored data, computed on the GPU through the C ABI.

Mirrors the reference functions (same names and argument meaning):
  comp_ac_fft           src/stats/summary.jl:46-63, 79-89
  comp_ac_time_missing  src/stats/summary.jl:455-496
  comp_psd_adfriendly   src/stats/summary.jl:183-235
  comp_psd              src/stats/summary.jl:111-165 (method = "periodogram": DSP.periodogram)
  comp_psd_lombscargle  src/stats/summary.jl:305-356
Arrays are (trials, time) with `dims` = last, the reference default; a Fortran-ordered
(column-major) array is passed through without a copy, exactly as a Julia Matrix would be.
"""
import numpy as np

from . import _lib


def _as_slices(data):
    data = np.asarray(data, dtype=np.float64)
    if data.ndim == 1:
        data = data.reshape(1, -1)
    if data.ndim != 2:
        raise ValueError("data must be a vector or a (trials, time) matrix")
    return np.asfortranarray(data)


def _strided(data):
    """(array, n_slices, n_t, slice_stride, time_stride) of a (trials, time) matrix as it lies in memory: a
    C-ordered NumPy matrix (= a Julia matrix with time down the columns, `dims = 1`) goes to the device
    untransposed (itjl_acf_strided)."""
    data = np.asarray(data, dtype=np.float64)
    if data.ndim == 1:
        data = data.reshape(1, -1)
    if data.ndim != 2:
        raise ValueError("data must be a vector or a (trials, time) matrix")
    if not (data.flags.c_contiguous or data.flags.f_contiguous) or data.size == 0:
        data = np.asfortranarray(data)
    es = data.itemsize
    return data, data.shape[0], data.shape[1], max(data.strides[0] // es, 1), max(data.strides[1] // es, 1)


def comp_ac_fft(data, n_lags=None, ctx=None):
    """Autocorrelation of each trial (rows), lags 0..n_lags-1; returns (trials, n_lags)
    (a vector input returns a vector, as the reference's Vector method does)."""
    vec = np.ndim(data) == 1
    d, n_slices, n_t, ss, ts = _strided(data)
    ctx = ctx or _lib.default_context()
    n_lags = n_t if n_lags is None else int(n_lags)
    out = np.empty((n_slices, n_lags), dtype=np.float64, order="F")
    ctx.check(_lib.lib().itjl_acf_strided(ctx.handle, _lib.ptr(d), n_slices, n_t, ss, ts, n_lags, 0, _lib.ptr(out)))
    return out[0].copy() if vec else out


def comp_ac_time_missing(data, n_lags=None, ctx=None):
    """NaN-aware autocorrelation (NaN = missing sample)."""
    vec = np.ndim(data) == 1
    d, n_slices, n_t, ss, ts = _strided(data)
    ctx = ctx or _lib.default_context()
    n_lags = n_t if n_lags is None else int(n_lags)
    out = np.empty((n_slices, n_lags), dtype=np.float64, order="F")
    ctx.check(_lib.lib().itjl_acf_strided(ctx.handle, _lib.ptr(d), n_slices, n_t, ss, ts, n_lags, 1, _lib.ptr(out)))
    return out[0].copy() if vec else out


def _ac_next_pow_two(i):
    """src/stats/bat_autocor.jl:21-24."""
    p = 1
    while p < i:
        p <<= 1
    return p


def comp_psd_adfriendly(data, fs, ctx=None):
    """Hamming-windowed, zero-padded periodogram; returns (power, freqs) with the DC bin
    dropped: power is (trials, n2/2 - 1), n2 = 2 * nextpow2(n)."""
    vec = np.ndim(data) == 1
    d = _as_slices(data)
    ctx = ctx or _lib.default_context()
    n_slices, n_t = d.shape
    n2 = 2 * _ac_next_pow_two(n_t)
    nb = n2 // 2 - 1
    out = np.empty((n_slices, nb), dtype=np.float64, order="F")
    ctx.check(_lib.lib().itjl_psd_hamming(ctx.handle, _lib.ptr(d), n_slices, n_t, float(fs), _lib.ptr(out)))
    freqs = np.arange(1, n2 // 2, dtype=np.float64) * (float(fs) / n2)     # fftfreq(n2, fs)[2:n2/2]
    return (out[0].copy() if vec else out), freqs


def nextfastfft(n):
    """DSP.nextfastfft(n) = nextprod([2, 3, 5, 7], n)."""
    return int(_lib.lib().itjl_nextfastfft(int(n)))


def comp_psd(data, fs, method="periodogram", ctx=None):
    """comp_psd(x, fs; method = "periodogram"): DSP.periodogram(x; fs, window = hamming) per trial with the DC
    bin dropped; returns (power, freqs), power (trials, nfft / 2), nfft = nextfastfft(n)."""
    if method != "periodogram":
        raise NotImplementedError('comp_psd: only method = "periodogram" is on the B200 path (the reference itself '
                                  "warns about its Welch defaults)")
    vec = np.ndim(data) == 1
    d = _as_slices(data)
    ctx = ctx or _lib.default_context()
    n_slices, n_t = d.shape
    nfft = nextfastfft(n_t)
    out = np.empty((n_slices, nfft // 2), dtype=np.float64, order="F")
    ctx.check(_lib.lib().itjl_psd_periodogram(ctx.handle, _lib.ptr(d), n_slices, n_t, float(fs), _lib.ptr(out)))
    freqs = np.arange(1, nfft // 2 + 1, dtype=np.float64) * (float(fs) / nfft)     # rfftfreq(nfft, fs)[2:end]
    return (out[0].copy() if vec else out), freqs


def lombscargle_autofrequency(times, maximum_frequency, samples_per_peak=5):
    """LombScargle.autofrequency(times; maximum_frequency): df = 1 / (samples_per_peak * (max - min)),
    grid df/2 : df : maximum_frequency."""
    times = np.asarray(times, dtype=np.float64)
    df = 1.0 / (samples_per_peak * (times.max() - times.min()))
    f0 = 0.5 * df
    n = int(np.floor((maximum_frequency - f0) / df + 1e-9)) + 1
    return f0, df, n


def comp_psd_lombscargle(times, data, nanmask=None, dt=None, ctx=None):
    """comp_psd_lombscargle(times, data, nanmask, dt; dims = 2): Lomb-Scargle periodogram of every trial's valid
    samples on LombScargle.autofrequency(times; maximum_frequency = 1 / 2dt); returns (power, frequency_grid).
    `times` must be the regular grid dt:dt:n*dt the reference's callers pass (acw.jl:247)."""
    vec = np.ndim(data) == 1
    d = _as_slices(data)                                   # no copy when the matrix already has the slice index fastest
    if nanmask is not None:
        m = np.asarray(nanmask, dtype=bool).reshape(d.shape)
        if not np.all(np.isnan(d[m])):                     # the reference's callers pass isnan.(data): nothing to mark then
            d = np.array(d, order="F", copy=True)
            d[m] = np.nan
    times = np.asarray(times, dtype=np.float64)
    dt = float(times[1] - times[0]) if dt is None else float(dt)
    if times.shape[0] != d.shape[1] or not np.allclose(np.diff(times), dt, rtol=1e-9, atol=0.0):
        raise ValueError("comp_psd_lombscargle: times must be a regular grid matching the data")
    ctx = ctx or _lib.default_context()
    f0, df, n_freq = lombscargle_autofrequency(times, 1.0 / (2.0 * dt))
    n_slices, n_t = d.shape
    out = np.empty((n_slices, n_freq), dtype=np.float64, order="F")
    ctx.check(_lib.lib().itjl_psd_lombscargle(ctx.handle, _lib.ptr(d), n_slices, n_t, dt, f0, df, n_freq, _lib.ptr(out)))
    freqs = f0 + df * np.arange(n_freq, dtype=np.float64)
    return (out[0].copy() if vec else out), freqs
