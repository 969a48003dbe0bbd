"""Host-side mirror of the reference's ACW helper functions (src/utils/utils.jl).

`acw0`, `acw50`, `acweuler` index a caller-provided `lags` vector with the first lag whose
ACF value is <= 0 / 0.5 / 1/e (utils.jl:257-264, 218-226, 294-302); on the batched `acw`
path those indices come from the GPU.  `fit_expdecay` runs the exp-decay fit kernel.
`find_knee_frequency`, `fooof_fit`, `find_oscillation_peak`, `lorentzian_initial_guess` (utils.jl:423-765) run the
batched knee kernel (itjl_fit_knee): the argmin of the reference's mean-absolute-deviation objective from the
reference's initial guess - see include/itjl_b200.h for what that does and does not pin.
"""
import math
import warnings

import numpy as np

from . import _lib

INV_E = 1.0 / math.e


def _first_leq(lags, acf, thr, what):
    acf = np.asarray(acf, dtype=np.float64)
    hit = np.nonzero(acf <= thr)[0]
    if hit.size == 0:
        warnings.warn(f"No {what} crossings found in ACF. Returning NaN.")
        return float("nan")
    return float(lags[int(hit[0])])


def acw0(lags, acf):
    return _first_leq(lags, acf, 0.0, "zero")


def acw50(lags, acf):
    return _first_leq(lags, acf, 0.5, "50%")


def acweuler(lags, acf):
    return _first_leq(lags, acf, INV_E, "1/e")


def tau_from_acw50(acw50_value):
    """utils.jl:196."""
    return -acw50_value / math.log(0.5)


def acw50_analytical(tau):
    """utils.jl:195."""
    return -tau * math.log(0.5)


def expdecay(tau, lags):
    """utils.jl:53-55."""
    return np.exp(-(1.0 / tau) * np.asarray(lags, dtype=np.float64))


def fit_expdecay(lags, acf, ctx=None, _entry="itjl_fit_expdecay"):
    """utils.jl:113-138: timescale of exp(-lags/tau) fitted to `acf` (vector -> scalar,
    (rows, lags) matrix -> vector).  `lags` must be uniformly spaced from 0."""
    lags = np.asarray(lags, dtype=np.float64)
    acf = np.asarray(acf, dtype=np.float64)
    vec = acf.ndim == 1
    rows = np.asfortranarray(acf.reshape(1, -1) if vec else acf)
    n_rows, n_lags = rows.shape
    if lags.shape[0] != n_lags:
        raise ValueError("lags and acf lengths differ")
    dt = float(lags[1] - lags[0]) if n_lags > 1 else 1.0
    if n_lags > 2 and not np.allclose(np.diff(lags), dt, rtol=1e-9, atol=0.0) or lags[0] != 0.0:
        raise ValueError("fit_expdecay on the GPU path needs lags = (0:n-1)*dt")
    ctx = ctx or _lib.default_context()
    out = np.empty(n_rows, dtype=np.float64)
    ctx.check(getattr(_lib.lib(), _entry)(ctx.handle, _lib.ptr(rows), n_rows, n_lags, dt, _lib.ptr(out)))
    return float(out[0]) if vec else out


def fit_expdecay_3_parameters(lags, acf, ctx=None):
    """utils.jl:159-192: tau of A (exp(-t / tau) + B) fitted to acf[2:end] over lags[2:end] (what
    acw(...; skip_zero_lag = true) uses); vector -> scalar, (rows, lags) matrix -> vector."""
    return fit_expdecay(lags, acf, ctx=ctx, _entry="itjl_fit_expdecay_3_parameters")


# ---- PSD-side reducers (utils.jl:197-198, 359-361, 423-765) ------------------------------------------
def tau_from_knee(knee):
    """utils.jl:197."""
    return 1.0 / (2.0 * np.pi * np.asarray(knee, dtype=np.float64))


def knee_from_tau(tau):
    """utils.jl:198."""
    return 1.0 / (2.0 * np.pi * np.asarray(tau, dtype=np.float64))


def lorentzian(f, u):
    """utils.jl:359-361: amp / (1 + (f / knee)^2)."""
    return u[0] / (1.0 + (np.asarray(f, dtype=np.float64) / u[1]) ** 2)


def _knee_call(psd, freqs, min_freq, max_freq, mode, max_peaks, ctx):
    psd = np.asarray(psd, dtype=np.float64)
    freqs = np.ascontiguousarray(freqs, dtype=np.float64)
    vec = psd.ndim == 1
    rows = np.ascontiguousarray(psd.reshape(1, -1) if vec else psd)
    if rows.ndim != 2 or rows.shape[1] != freqs.shape[0]:
        raise ValueError("psd must be a vector or a (rows, freqs) matrix matching freqs")
    min_freq = float(freqs[0]) if min_freq is None else float(min_freq)
    max_freq = float(freqs[-1]) if max_freq is None else float(max_freq)
    ctx = ctx or _lib.default_context()
    n_rows, n_freq = rows.shape
    knee = np.empty(n_rows); amp = np.empty(n_rows); peak = np.empty(n_rows)
    ctx.check(_lib.lib().itjl_fit_knee(ctx.handle, _lib.ptr(rows), n_rows, n_freq, 0, _lib.ptr(freqs), min_freq, max_freq,
                                       int(mode), int(max_peaks), _lib.ptr(knee), _lib.ptr(amp), _lib.ptr(peak)))
    return vec, amp, knee, peak


def lorentzian_initial_guess(psd, freqs, min_freq=None, max_freq=None, ctx=None):
    """utils.jl:423-445 -> [amp, knee] (vector) or (rows, 2)."""
    vec, amp, knee, _ = _knee_call(psd, freqs, min_freq, max_freq, 2, 0, ctx)
    return [float(amp[0]), float(knee[0])] if vec else np.stack([amp, knee], axis=1)


def find_knee_frequency(psd, freqs, min_freq=None, max_freq=None, allow_variable_exponent=False, constrained=False,
                        ctx=None, return_peak=False):
    """utils.jl:479-538 -> [amp, knee] for a vector, (rows, 2) for a (rows, freqs) matrix.  return_peak adds
    find_oscillation_peak(psd - lorentzian(freqs, [amp, knee]), freqs; min_freq, max_freq) of every row."""
    if allow_variable_exponent or constrained:
        raise NotImplementedError("only constrained = false, allow_variable_exponent = false is on the B200 path")
    vec, amp, knee, peak = _knee_call(psd, freqs, min_freq, max_freq, 0, 0, ctx)
    out = [float(amp[0]), float(knee[0])] if vec else np.stack([amp, knee], axis=1)
    if return_peak:
        return out, (float(peak[0]) if vec else peak)
    return out


def find_oscillation_peak(psd, freqs, min_freq=5.0 / 1000.0, max_freq=50.0 / 1000.0, ctx=None):
    """utils.jl:721-765 (min_prominence_ratio = 0.1): frequency of the most prominent peak, NaN if none."""
    vec, _, _, peak = _knee_call(psd, freqs, min_freq, max_freq, 3, 0, ctx)
    return float(peak[0]) if vec else peak


def fooof_fit(psd, freqs, min_freq=None, max_freq=None, oscillation_peak=True, max_peaks=3, return_only_knee=False,
              allow_variable_exponent=False, constrained=False, ctx=None):
    """utils.jl:598-675.  Returns the knee frequency (return_only_knee = true, what acw uses) or, as the reference
    does without it, (knee, first oscillation peak frequency) - the Gaussian amplitudes / widths stay on the device."""
    if allow_variable_exponent or constrained:
        raise NotImplementedError("only constrained = false, allow_variable_exponent = false is on the B200 path")
    freqs = np.asarray(freqs, dtype=np.float64)
    min_freq = float(freqs[0]) if min_freq is None else float(min_freq)
    max_freq = float(freqs[-1]) if max_freq is None else float(max_freq)
    if oscillation_peak:
        vec, _, knee, peak = _knee_call(psd, freqs, min_freq, max_freq, 1, max_peaks, ctx)
    else:
        m = (freqs >= min_freq) & (freqs <= max_freq)                                  # utils.jl:609-611
        vec, _, knee, peak = _knee_call(np.asarray(psd, dtype=np.float64)[..., m], freqs[m], min_freq, max_freq, 0, 0, ctx)
    k = float(knee[0]) if vec else knee
    if return_only_knee or not oscillation_peak:
        return k
    return k, (float(peak[0]) if vec else peak)
