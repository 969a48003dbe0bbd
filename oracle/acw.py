"""Oracle: the `acw()` front end on (trials, time) data.

Follows /root/reference/src/core/acw.jl:109-307 (dims = last, trial_dims = 1):
  complete data -> comp_ac_fft, all n lags           :145-146
  NaN data      -> comp_ac_time_missing[, n_lags]    :147-154
  average_over_trials                                :156-158
  acw0_sample (sample-index lags)                    :160-163
  :acw0 / :acw50 / :acweuler                         :164-191
  :auc window = floor(Int, acw0_sample[1]) of the FIRST slice   :192-201
  n_lags = ceil(Int, 1.1 * nanmaximum(acw0_sample)) or n        :203-210
  :tau on acf[:, 1:n_lags]                           :212-231
  :knee: comp_psd (complete) / comp_psd_lombscargle on dt:dt:n dt (NaN data), mean over trials,
         tau_from_knee(fooof_fit(...; min_freq, max_freq = freqlims, return_only_knee = true))   :242-270
"""
import math

import numpy as np

from . import acwred, knee, summary

ACF_ACWTYPES = ("acw0", "acw50", "acweuler", "auc", "tau")


def acw_knee(data, fs, freqlims=None, average_over_trials=False, oscillation_peak=True, max_peaks=1):
    """acw.jl:242-270 on (trials, time) data -> (tau per slice, psd, freqs)."""
    data = np.asarray(data, dtype=np.float64)
    if data.ndim == 1:
        data = data.reshape(1, -1)
    dt = 1.0 / fs
    mask = np.isnan(data)
    if mask.any():
        psd, freqs = summary.comp_psd_lombscargle(dt * np.arange(1, data.shape[1] + 1), data, mask, dt)
    else:
        psd, freqs = summary.comp_psd_periodogram(data, fs)
    if average_over_trials:
        psd = psd.mean(axis=0, keepdims=True)
    lo, hi = (freqs[0], freqs[-1]) if freqlims is None else freqlims
    tau = np.array([knee.tau_from_knee(knee.fooof_fit(p, freqs, lo, hi, oscillation_peak, max_peaks, True)) for p in psd])
    return tau, psd, freqs


def acw(data, fs, acwtypes=ACF_ACWTYPES, n_lags=None, average_over_trials=False, skip_zero_lag=False, freqlims=None,
        oscillation_peak=True, max_peaks=1):
    data = np.asarray(data, dtype=np.float64)
    if data.ndim == 1:
        data = data.reshape(1, -1)                      # acw.jl:122-125
    if isinstance(acwtypes, str):
        acwtypes = (acwtypes,)
    for t in acwtypes:
        if t not in ACF_ACWTYPES + ("knee",):
            raise ValueError(f"unsupported acwtype {t!r}")
    if "knee" in acwtypes:
        tau, psd, freqs = acw_knee(data, fs, freqlims, average_over_trials, oscillation_peak, max_peaks)
        if tuple(acwtypes) == ("knee",):
            return {"knee": tau, "psd": psd, "freqs": freqs}
        rest = acw(data, fs, [t for t in acwtypes if t != "knee"], n_lags, average_over_trials, skip_zero_lag)
        rest.update(knee=tau, psd=psd, freqs=freqs)
        return rest
    dt = 1.0 / fs
    n = data.shape[1]
    iscomplete = not np.isnan(data).any()
    if iscomplete:
        acf = summary.comp_ac_fft(data)
    else:
        acf = summary.comp_ac_time_missing(data, n_lags)
    if average_over_trials:
        acf = acf.mean(axis=0, keepdims=True)
    n_slices = acf.shape[0]
    lags = np.arange(n, dtype=np.float64) * dt

    k0 = np.array([acwred.first_crossing_index(a, 0.0) for a in acf])
    k50 = np.array([acwred.first_crossing_index(a, 0.5) for a in acf])
    ke = np.array([acwred.first_crossing_index(a, acwred.INV_E) for a in acf])

    def as_lag(k):
        return np.where(k >= 0, k * dt, np.nan)

    out = {"k0": k0, "k50": k50, "ke": ke}
    if "acw0" in acwtypes:
        out["acw0"] = as_lag(k0)
    if "acw50" in acwtypes:
        out["acw50"] = as_lag(k50)
    if "acweuler" in acwtypes:
        out["acweuler"] = as_lag(ke)
    if "auc" in acwtypes:
        if k0[0] < 0:
            raise ValueError("InexactError: Int64(NaN)")          # floor(Int, NaN)
        w = int(k0[0])
        out["auc"] = np.array([acwred.acw_romberg(dt, a[:w]) for a in acf])
    if n_lags is None:
        valid = k0[k0 >= 0]
        if valid.size == 0:
            n_lags = n
        else:
            n_lags = int(math.ceil(1.1 * float(valid.max())))
    acf = acf[:, :n_lags]
    lags = lags[:n_lags]
    if "tau" in acwtypes:
        fit = acwred.fit_expdecay_3_parameters if skip_zero_lag else acwred.fit_expdecay       # acw.jl:212-220
        out["tau"] = np.array([fit(lags, a) for a in acf])
    out["n_lags"] = n_lags
    out["acf"] = acf
    out["lags"] = lags
    out["n_slices"] = n_slices
    return out
