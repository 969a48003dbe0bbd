"""`acw(data, fs; ...)` on the GPU - mirror of src/core/acw.jl:109-307 for the ACF-based
acwtypes (:acw0, :acw50, :acweuler, :auc, :tau) and for :knee (acw.jl:242-270: periodogram or Lomb-Scargle
spectrum of every slice, then the FOOOF-style knee fit - itjl_acw_knee).
"""
import ctypes
import math
from dataclasses import dataclass
from typing import Any

import numpy as np

from . import _lib

possible_acwtypes = ["acw0", "acw50", "acweuler", "auc", "tau", "knee"]
_BITS = {"acw0": _lib.ACW0, "acw50": _lib.ACW50, "acweuler": _lib.ACWEULER,
         "auc": _lib.AUC, "tau": _lib.TAU}


@dataclass
class ACWResults:
    """src/core/acw.jl:44-55."""
    fs: float
    acw_results: Any
    acwtypes: Any
    n_lags: Any
    freqlims: Any
    acf: Any
    psd: Any
    freqs: Any
    lags: Any
    x_dim: Any


class DeviceData:
    """A (slices x time) matrix kept on the device by `upload(data, dims=...)`: repeated acw() calls on it
    (different acwtypes, n_lags, averaging) pay no host -> device copy (itjl_data_upload / itjl_acw_resident).
    Any other data call on the same context replaces it."""

    def __init__(self, ctx, n_slices, n_t, other_shape, t_axis, was_vector):
        self.ctx, self.n_slices, self.n_t = ctx, n_slices, n_t
        self.other_shape, self.t_axis, self.was_vector = other_shape, t_axis, was_vector


def _matrix_view(data, dims):
    """(array to pass, n_slices, n_t, slice_stride, time_stride, other_shape, t_axis, was_vector): a 2-D input is
    handed to the library as it lies in memory with its element strides (the device re-lays it: no host
    transpose); N-d inputs are flattened to (slices x time) on the host in the reference's slice order."""
    data = np.asarray(data, dtype=np.float64)
    was_vector = data.ndim == 1
    if was_vector:
        data = data.reshape(1, -1)                              # acw.jl:122-125
        dims = None
    nd = data.ndim
    if dims is None or dims == -1:
        dims = nd
    if not (isinstance(dims, (int, np.integer)) and 1 <= dims <= nd):
        raise ValueError(f"dims must be a dimension of data (1 .. {nd}, or -1 for the last one), got {dims!r}")
    t_axis = int(dims) - 1
    other_shape = tuple(n for i, n in enumerate(data.shape) if i != t_axis)
    if nd == 2 and (data.flags.c_contiguous or data.flags.f_contiguous) and data.size and min(data.shape) > 1:
        s_axis = 1 - t_axis
        es = data.itemsize
        return (data, data.shape[s_axis], data.shape[t_axis], data.strides[s_axis] // es, data.strides[t_axis] // es,
                other_shape, t_axis, was_vector)
    # slices in the reference's order (column-major over the remaining dimensions), time last
    d = np.asfortranarray(np.moveaxis(data, t_axis, -1).reshape((-1, data.shape[t_axis]), order="F"))
    return d, d.shape[0], d.shape[1], 1, d.shape[0], other_shape, t_axis, was_vector


def upload(data, dims=None, ctx=None):
    """Put `data` on the device once; pass the returned DeviceData to acw() as often as needed."""
    d, n_slices, n_t, ss, ts, other_shape, t_axis, was_vector = _matrix_view(data, dims)
    ctx = ctx or _lib.default_context()
    ctx.check(_lib.lib().itjl_data_upload(ctx.handle, _lib.ptr(d), n_slices, n_t, int(ss), int(ts)))
    return DeviceData(ctx, n_slices, n_t, other_shape, t_axis, was_vector)


def _acw_knee(data, dims, fs, freqlims, average_over_trials, oscillation_peak, max_peaks, return_psd, ctx):
    """The :knee branch (acw.jl:242-270) -> (tau per slice, psd, freqs, freqlims)."""
    from .utils import tau_from_knee
    d, n_slices, n_t, ss, ts, other_shape, t_axis, was_vector = _matrix_view(data, dims)
    if not (ss == 1 and ts == n_slices):
        # (slices x time) with the slice index fastest, as the library's spectrum kernels read it
        d = np.asfortranarray(np.moveaxis(np.asarray(data, dtype=np.float64).reshape(
            (1, -1) if was_vector else np.shape(data)), t_axis, -1).reshape((-1, n_t), order="F"))
    if average_over_trials and len(other_shape) != 1:
        raise NotImplementedError("average_over_trials is implemented for (trials, time) matrices")
    has_nan = bool(np.isnan(d).any())
    nb = int(_lib.lib().itjl_acw_knee_nfreq(n_t, int(has_nan)))
    if nb < 3:
        raise ValueError("acw(:knee): the time series is too short or too long for the spectrum kernels")
    n_out = 1 if average_over_trials else n_slices
    knee = np.empty(n_out, dtype=np.float64)
    psd = np.empty((n_out, nb), dtype=np.float64, order="F") if return_psd else None
    freqs = np.empty(nb, dtype=np.float64)
    lo, hi = (float("nan"), float("nan")) if freqlims is None else (float(freqlims[0]), float(freqlims[1]))
    nf = ctypes.c_int64(0)
    flags = (1 if oscillation_peak else 0) | (2 if average_over_trials else 0)
    ctx.check(_lib.lib().itjl_acw_knee(ctx.handle, _lib.ptr(d), n_slices, n_t, float(fs), lo, hi, flags, int(max_peaks),
                                       _lib.ptr(knee), _lib.ptr(psd), 0 if psd is None else psd.size, _lib.ptr(freqs),
                                       ctypes.byref(nf)))
    if freqlims is None:
        freqlims = (float(freqs[0]), float(freqs[-1]))                      # acw.jl:255-257
    tau = tau_from_knee(knee)
    if was_vector or n_out == 1:
        tau_out = float(tau[0])
    else:
        tau_out = tau.reshape(other_shape, order="F")
    if psd is not None:
        if was_vector:
            psd = psd[0]
        elif n_out == n_slices:
            psd = np.moveaxis(psd.reshape(other_shape + (nb,), order="F"), -1, t_axis)
    return tau_out, psd, freqs, freqlims


def acw(data, fs, acwtypes=None, n_lags=None, dims=None, return_acf=True,
        average_over_trials=False, ctx=None, freqlims=None, return_psd=True, oscillation_peak=True, max_peaks=1,
        allow_variable_exponent=False, constrained=False, skip_zero_lag=False):
    """Compute ACW measures of every slice of `data` along its time dimension.

    `dims` is the time dimension in the reference's convention: 1-based, default = the last one
    (acw.jl:113); -1 also means last.  Slices run over all other dimensions (get_slices,
    utils.jl:863-866); results have the shape of `data` without the time dimension and `acf` has the lags
    where time was (stack_and_reshape, utils.jl:890-895).  The library call itself always sees a
    (slices x time) matrix with the slice index fastest.

    Returns ACWResults; `acw_results` is a list in the order of `acwtypes` (a single entry when one type
    is requested), each entry an array over slices (scalar for a vector input or
    average_over_trials=True).  Lags without a crossing give NaN.
    """
    if acwtypes is None:
        acwtypes = list(possible_acwtypes)
    single = isinstance(acwtypes, str)
    if single:
        acwtypes = [acwtypes]
    acwtypes = [str(t).lstrip(":") for t in acwtypes]
    if len(acwtypes) == 0:
        raise ValueError(f"No ACW types specified. Possible ACW types: {possible_acwtypes}")
    for t in acwtypes:
        if t not in _BITS and t != "knee":
            raise ValueError(f"Possible ACW types: {possible_acwtypes}, got {t!r}")
    if allow_variable_exponent or constrained:
        raise NotImplementedError("only constrained = false, allow_variable_exponent = false is on the B200 path")
    resident = isinstance(data, DeviceData)
    knee_out = None
    if "knee" in acwtypes:
        if resident:
            raise NotImplementedError("acw(:knee) needs the host array (the spectrum kernels upload it themselves)")
        knee_out = _acw_knee(data, dims, fs, freqlims, average_over_trials, oscillation_peak, max_peaks, return_psd,
                             ctx or _lib.default_context())
        if all(t == "knee" for t in acwtypes):
            tau_k, psd, freqs, fl = knee_out
            return ACWResults(float(fs), tau_k if (single or len(acwtypes) == 1) else [tau_k] * len(acwtypes),
                              acwtypes[0] if single else acwtypes, None, fl, None, psd, freqs, None,
                              None if psd is None else psd.ndim)
    if resident:
        ctx = data.ctx
        d, n_slices, n_t, ss, ts = None, data.n_slices, data.n_t, 1, data.n_slices
        other_shape, t_axis, was_vector = data.other_shape, data.t_axis, data.was_vector
    else:
        d, n_slices, n_t, ss, ts, other_shape, t_axis, was_vector = _matrix_view(data, dims)
        ctx = ctx or _lib.default_context()
    if average_over_trials and len(other_shape) != 1:
        raise NotImplementedError("average_over_trials is implemented for (trials, time) matrices")
    mask = 0
    for t in acwtypes:
        mask |= _BITS.get(t, 0)
    if skip_zero_lag and (mask & _lib.TAU):
        mask = (mask & ~_lib.TAU) | _lib.TAU3              # fit_expdecay_3_parameters over lags[2:end] (acw.jl:212-216)
    n_out = 1 if average_over_trials else n_slices
    k0 = np.empty(n_out, dtype=np.int32)
    k50 = np.empty(n_out, dtype=np.int32)
    ke = np.empty(n_out, dtype=np.int32)
    auc = np.empty(n_out, dtype=np.float64)
    tau = np.empty(n_out, dtype=np.float64)
    nl = ctypes.c_int64(0 if n_lags is None else int(n_lags))
    # n_lags is only known after the call: the ACF stays on the device and is fetched below into an array of
    # exactly (n_out, n_lags) (itjl_acw_acf) - a worst-case (n_out x n_t) buffer costs a page fault per 4 KB written
    acf_buf = None
    cap = 0
    # always request acw0: the AUC window and the default n_lags derive from it
    if resident:
        rc = _lib.lib().itjl_acw_resident(ctx.handle, float(fs), mask | _lib.ACW0, int(bool(average_over_trials)),
                                          ctypes.byref(nl), _lib.ptr(k0), _lib.ptr(k50), _lib.ptr(ke), _lib.ptr(auc),
                                          _lib.ptr(tau), _lib.ptr(acf_buf), cap)
    else:
        rc = _lib.lib().itjl_acw_strided(ctx.handle, _lib.ptr(d), n_slices, n_t, int(ss), int(ts),
                                         float(fs), mask | _lib.ACW0, int(bool(average_over_trials)), ctypes.byref(nl),
                                         _lib.ptr(k0), _lib.ptr(k50), _lib.ptr(ke), _lib.ptr(auc), _lib.ptr(tau),
                                         _lib.ptr(acf_buf), cap)
    ctx.check(rc)
    n_lags_used = int(nl.value)
    dt = 1.0 / float(fs)
    if return_acf:
        acf_buf = np.empty(n_out * n_lags_used, dtype=np.float64)
        ctx.check(_lib.lib().itjl_acw_acf(ctx.handle, _lib.ptr(acf_buf), acf_buf.size, None, None))

    def as_lag(k):
        return np.where(k >= 0, k.astype(np.float64) * dt, np.nan)

    if "auc" in acwtypes:
        # reference: floor(Int, NaN) -> InexactError; romberg on < 2 points -> ArgumentError
        if k0[0] < 0:
            raise ValueError("InexactError: no zero crossing in the first slice (AUC window undefined)")
        if k0[0] < 2:
            raise ValueError("ArgumentError: y array must have length > 1 (AUC window of the first slice)")
    res = {"acw0": as_lag(k0), "acw50": as_lag(k50), "acweuler": as_lag(ke), "auc": auc, "tau": tau}
    results = []
    for t in acwtypes:
        if t == "knee":
            results.append(knee_out[0])
            continue
        v = res[t]
        if was_vector or n_out == 1:
            results.append(float(v[0]))
        else:
            results.append(v.reshape(other_shape, order="F"))
    acf = lags = None
    if return_acf:
        acf = acf_buf[: n_out * n_lags_used].reshape((n_out, n_lags_used), order="F")
        if was_vector:
            acf = acf[0]
        elif n_out == n_slices:
            # lags go where time was
            acf = np.moveaxis(acf.reshape(other_shape + (n_lags_used,), order="F"), -1, t_axis)
        lags = np.arange(n_lags_used, dtype=np.float64) * dt
    x_dim = None if acf is None else acf.ndim
    psd, freqs, fl = (knee_out[1], knee_out[2], knee_out[3]) if knee_out is not None else (None, None, None)
    return ACWResults(float(fs), results[0] if len(results) == 1 else results,
                      acwtypes[0] if single else acwtypes, n_lags_used, fl, acf, psd, freqs,
                      lags, x_dim)
