"""Host-side mirror of the reference's ACW helper functions (src/utils/utils.jl).

`acw0`, `acw50`, `acweuler` index a caller-provided `lags` vector with the first lag whose
ACF value is <= 0 / 0.5 / 1/e (utils.jl:257-264, 218-226, 294-302); on the batched `acw`
path those indices come from the GPU.  `fit_expdecay` runs the exp-decay fit kernel.
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


def fit_expdecay(lags, acf, ctx=None):
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
    ctx.check(_lib.lib().itjl_fit_expdecay(ctx.handle, _lib.ptr(rows), n_rows, n_lags, dt, _lib.ptr(out)))
    return float(out[0]) if vec else out
