"""Oracle: Ornstein-Uhlenbeck generators.

Follows /root/reference/src/utils/ou_process.jl:
  generate_ou_process          :46-62
  generate_ou_process_sciml    :114-148
  generate_ou_with_oscillation :176-218

The reference integrates  du = -u/tau dt + dW  (unit noise intensity,
ou_process.jl:64-67) from u(0) ~ N(0,1) (:126-131) with the adaptive SOSRA solver
and saves at t = dt, 2dt, ..., T (:133).  To solver tolerance that is the exact OU
transition  x_{k+1} = a x_k + sqrt(tau/2 (1 - a^2)) xi_k,  a = exp(-dt/tau); that
transition ("exact") is what the oracle and the CUDA path implement, with an
Euler-Maruyama variant ("em": a = 1 - dt/tau, s = sqrt(dt)) selectable.
SOSRA's internal noise stream cannot be reproduced => path-wise parity with the
Julia solver is unpinned; the distributional tests of test/test_ou_process.jl:18-31
pin it instead.
"""
import numpy as np

from . import philox


def n_time_points(dt, duration):
    """length(dt:dt:duration)  (ou_process.jl:133)."""
    return int(np.floor(duration / dt + 1e-9))


def ou_coefficients(tau, dt, update="exact"):
    """AR(1) coefficient a and innovation scale s of one dt step."""
    tau = float(tau)
    if update == "exact":
        with np.errstate(over="ignore", invalid="ignore", divide="ignore"):
            a = np.exp(-dt / tau)
            s = np.sqrt(tau / 2.0 * (1.0 - a * a))
    elif update == "em":
        a = 1.0 - dt / tau
        s = np.sqrt(dt)
    else:
        raise ValueError(update)
    return float(a), float(s)


def ou_raw(tau, dt, n_t, u0, noise, update="exact"):
    """Unstandardised trajectories.  u0: (trials,), noise: (trials, n_t).
    Returns (trials, n_t); sample k is the state at t = (k+1) dt."""
    a, s = ou_coefficients(tau, dt, update)
    u0 = np.asarray(u0, dtype=np.float64)
    noise = np.asarray(noise, dtype=np.float64)
    x = np.empty_like(noise)
    prev = u0.copy()
    with np.errstate(over="ignore", invalid="ignore"):
        for k in range(n_t):
            prev = a * prev + s * noise[:, k]
            x[:, k] = prev
    return x


def standardize_rows(x, scale):
    """((x - mean_t) / std_t(n-1)) * scale per trial  (ou_process.jl:143)."""
    with np.errstate(over="ignore", invalid="ignore", divide="ignore"):
        m = x.mean(axis=1, keepdims=True)
        sd = x.std(axis=1, ddof=1, keepdims=True)
        return (x - m) / sd * scale


def draw_inputs(seed, generation, particle, n_trials, n_t):
    """Philox-drawn (u0, phases, noise) for one particle (see philox.py)."""
    u0 = np.empty(n_trials)
    ph = np.empty(n_trials)
    noise = np.empty((n_trials, n_t))
    for tr in range(n_trials):
        u0[tr], ph[tr] = philox.init_state_and_phase(seed, generation, particle, tr)
        noise[tr] = philox.noise_normals(seed, generation, particle, tr, n_t)
    return u0, ph, noise


def generate_ou_process(tau, true_D, dt, duration, num_trials, standardize=True,
                        u0=None, noise=None, seed=0, generation=0, particle=0,
                        update="exact", n_t=None):
    """ou_process.jl:46-62 / :114-148.  Returns (num_trials, n_t) float64.
    Non-finite trajectories become all-NaN (the reference's solver-failure
    contract, :56-61)."""
    if n_t is None:
        n_t = n_time_points(dt, duration)
    if u0 is None or noise is None:
        du0, _, dnoise = draw_inputs(seed, generation, particle, num_trials, n_t)
        u0 = du0 if u0 is None else u0
        noise = dnoise if noise is None else noise
    x = ou_raw(tau, dt, n_t, u0, noise, update)
    if standardize:
        x = standardize_rows(x, true_D)
    if not np.all(np.isfinite(x)):
        x = np.full((num_trials, n_t), np.nan)
    return x


def clamp_coeff(coeff):
    """ou_process.jl:189-196: only values outside [0, 1] are moved."""
    if coeff < 0.0:
        return 1e-4
    if coeff > 1.0:
        return 1.0 - 1e-4
    return coeff


def generate_ou_with_oscillation(theta, dt, duration, num_trials, data_mean, data_sd,
                                 u0=None, noise=None, phases=None, seed=0,
                                 generation=0, particle=0, update="exact", n_t=None):
    """ou_process.jl:176-218."""
    tau, freq, coeff = float(theta[0]), float(theta[1]), float(theta[2])
    coeff = clamp_coeff(coeff)
    if n_t is None:
        n_t = n_time_points(dt, duration)
    if u0 is None or noise is None or phases is None:
        du0, dph, dnoise = draw_inputs(seed, generation, particle, num_trials, n_t)
        u0 = du0 if u0 is None else u0
        noise = dnoise if noise is None else noise
        phases = dph if phases is None else phases
    ou = ou_raw(tau, dt, n_t, u0, noise, update)              # standardize=false (:199)
    t = dt * np.arange(1, n_t + 1)                             # dt:dt:duration (:205)
    phases = np.asarray(phases, dtype=np.float64).reshape(-1, 1)
    oscil = np.sqrt(2.0) * np.sin(phases + 2.0 * np.pi * freq * t[None, :])   # :208
    with np.errstate(over="ignore", invalid="ignore", divide="ignore"):
        data = np.sqrt(1.0 - coeff) * oscil + np.sqrt(coeff) * ou             # :209
        data = standardize_rows(data, 1.0)                                    # :211
        out = data_sd * data + data_mean                                      # :213
    if not np.all(np.isfinite(out)):
        out = np.full((num_trials, n_t), np.nan)
    return out


def generate_two_timescale(theta, dt, duration, num_trials, data_var=1.0, seed=0, update="exact", n_t=None):
    """Models.generate_data(::TwoTimescaleModel, theta) (src/models/two_timescale.jl:30-51): the reference's unfinished
    two-timescale model only has this generator.  ou_i = generate_ou_process(tau_i, 1 / tau_i, ...) (standardised, times
    D_i = 1 / tau_i), mixed sqrt(coeff) : sqrt(1 - coeff), times sqrt(data_var).  Streams: particles 0 and 1 of `seed`."""
    tau1, tau2, coeff = theta
    o1 = generate_ou_process(tau1, 1.0 / tau1, dt, duration, num_trials, seed=seed, particle=0, update=update, n_t=n_t)
    o2 = generate_ou_process(tau2, 1.0 / tau2, dt, duration, num_trials, seed=seed, particle=1, update=update, n_t=n_t)
    return np.sqrt(data_var) * (np.sqrt(coeff) * o1 + np.sqrt(1.0 - coeff) * o2)
