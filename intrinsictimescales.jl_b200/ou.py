"""Ornstein-Uhlenbeck generators on the GPU.

Mirrors src/utils/ou_process.jl: generate_ou_process (:46-62), generate_ou_with_oscillation
(:176-218).  The reference solves the SDE with the adaptive SOSRA scheme and samples it
every dt; the kernels use the exact OU transition that solve converges to
(`update="exact"`, default) or Euler-Maruyama (`update="em"`).  Randomness comes from a
Philox4x32-7 stream keyed by `seed` (the reference's `rng` / `deq_seed` pair); `u0`,
`noise` and `phases` can be injected for parity tests.
"""
import math

import numpy as np

from . import _lib

_UPDATE = {"exact": _lib.UPDATE_EXACT, "em": _lib.UPDATE_EM}


def n_time_points(dt, duration):
    """length(dt:dt:duration) (ou_process.jl:133)."""
    return int(math.floor(duration / dt + 1e-9))


def _opt(a, shape):
    if a is None:
        return None
    a = np.ascontiguousarray(a, dtype=np.float64)
    if a.shape != shape:
        raise ValueError(f"expected shape {shape}, got {a.shape}")
    return a


def generate_ou_process(tau, true_D, dt, duration, num_trials, standardize=True, seed=0,
                        update="exact", u0=None, noise=None, ctx=None):
    """Returns a (num_trials, n_t) Float64 matrix (column-major, like the reference)."""
    ctx = ctx or _lib.default_context()
    n_t = n_time_points(dt, duration)
    num_trials = int(num_trials)
    out = np.empty((num_trials, n_t), dtype=np.float64, order="F")
    u0 = _opt(u0, (num_trials,))
    noise = _opt(noise, (num_trials, n_t))
    ctx.check(_lib.lib().itjl_generate_ou(ctx.handle, float(tau), float(true_D), float(dt), n_t,
                                          num_trials, int(bool(standardize)), _UPDATE[update],
                                          int(seed), _lib.ptr(u0), _lib.ptr(noise), _lib.ptr(out)))
    return out


def generate_ou_with_oscillation(theta, dt, duration, num_trials, data_mean, data_sd, seed=0,
                                 update="exact", u0=None, noise=None, phases=None, ctx=None):
    """theta = [tau, freq, coeff]."""
    ctx = ctx or _lib.default_context()
    n_t = n_time_points(dt, duration)
    num_trials = int(num_trials)
    out = np.empty((num_trials, n_t), dtype=np.float64, order="F")
    u0 = _opt(u0, (num_trials,))
    noise = _opt(noise, (num_trials, n_t))
    phases = _opt(phases, (num_trials,))
    ctx.check(_lib.lib().itjl_generate_ou_osc(ctx.handle, float(theta[0]), float(theta[1]),
                                              float(theta[2]), float(dt), n_t, num_trials,
                                              float(data_mean), float(data_sd), _UPDATE[update],
                                              int(seed), _lib.ptr(u0), _lib.ptr(noise),
                                              _lib.ptr(phases), _lib.ptr(out)))
    return out


def generate_two_timescale(theta, dt, duration, num_trials, data_var=1.0, seed=0, update="exact", ctx=None):
    """Models.generate_data(::TwoTimescaleModel, theta = [tau1, tau2, coeff]) (src/models/two_timescale.jl:30-51):
    sqrt(data_var) * (sqrt(coeff) * ou1 + sqrt(1 - coeff) * ou2), ou_i = generate_ou_process(tau_i, 1 / tau_i, ...)."""
    ctx = ctx or _lib.default_context()
    n_t = n_time_points(dt, duration)
    num_trials = int(num_trials)
    out = np.empty((num_trials, n_t), dtype=np.float64, order="F")
    ctx.check(_lib.lib().itjl_generate_ou_two(ctx.handle, float(theta[0]), float(theta[1]), float(theta[2]), float(dt), n_t,
                                              num_trials, float(data_var), _UPDATE[update], int(seed), _lib.ptr(out)))
    return out
