"""Model constructors and the model protocol, GPU-backed.

Mirrors (same names, argument meaning, defaults and error behaviour):
  one_timescale_model               src/models/one_timescale.jl:135-205 (ACF branch)
  one_timescale_with_missing_model  src/models/one_timescale_with_missing.jl:137-213 (ACF branch)
  one_timescale_and_osc_model       src/models/one_timescale_and_osc.jl:98-189
  check_model_inputs                src/core/model.jl:239-292
  draw_theta / generate_data / summary_stats / distance_function /
  generate_data_and_reduce          src/core/model.jl:146-156 + the model files
Every numeric step (data ACF/PSD, exp-decay prior fit, simulation, summary, distance) runs
through libitjl_b200; the per-particle protocol functions are batch-of-one calls of the same
kernels `basic_abc` uses.
"""
import math
from dataclasses import dataclass, field
from typing import Any

import numpy as np

from . import _lib, summary, utils
from .ou import n_time_points


# ---------------------------------------------------------------- priors ----
class Normal:
    """Distributions.Normal(mu, sigma)."""

    def __init__(self, mu, sigma):
        self.mu, self.sigma = float(mu), float(sigma)

    def rand(self, rng, size=None):
        return rng.normal(self.mu, self.sigma, size)

    def pdf(self, x):
        z = (np.asarray(x, dtype=np.float64) - self.mu) / self.sigma
        return np.exp(-0.5 * z * z) / (self.sigma * math.sqrt(2.0 * math.pi))

    def __repr__(self):
        return f"Normal(mu={self.mu}, sigma={self.sigma})"


class Uniform:
    """Distributions.Uniform(a, b)."""

    def __init__(self, a, b):
        self.a, self.b = float(a), float(b)

    def rand(self, rng, size=None):
        return rng.uniform(self.a, self.b, size)

    def pdf(self, x):
        x = np.asarray(x, dtype=np.float64)
        return np.where((x >= self.a) & (x <= self.b), 1.0 / (self.b - self.a), 0.0)

    def __repr__(self):
        return f"Uniform(a={self.a}, b={self.b})"


# ----------------------------------------------------------------- models ----
@dataclass
class TimescaleModel:
    """Field-compatible with OneTimescaleModel / OneTimescaleWithMissingModel /
    OneTimescaleAndOscModel (one_timescale.jl:61-83 and siblings)."""
    kind: str
    data: np.ndarray
    time: np.ndarray
    fit_method: str
    summary_method: str
    lags_freqs: np.ndarray
    prior: list
    distance_method: str
    data_sum_stats: np.ndarray
    dt: float
    T: float
    numTrials: int
    data_mean: float
    data_sd: float
    freqlims: Any = None
    n_lags: Any = None
    freq_idx: Any = None
    dims: int = 2
    distance_combined: bool = False
    weights: Any = None            # [w_stats, w_tau] of combined_distance (one_timescale.jl:278-290)
    data_tau: Any = None           # fit_expdecay(lags, data_sum_stats) / tau_from_knee(...) when distance_combined
    data_osc: Any = None           # oscillation peak of the data spectrum (oscillation models, PSD, distance_combined)
    missing_mask: Any = None
    update: str = "exact"
    _gpu: dict = field(default_factory=dict, repr=False)

    @property
    def n_t(self):
        return n_time_points(self.dt, self.T)

    @property
    def n_theta(self):
        return len(self.prior)

    # ---- GPU handle (itjl_model), cached per context --------------------------------
    def gpu_model(self, ctx=None):
        ctx = ctx or _lib.default_context()
        key = id(ctx)
        hit = self._gpu.get(key)
        if hit is None or hit.ctx is not ctx or not hit.handle:      # (an id can be reused by a later context)
            hit = self._gpu[key] = GpuModel(self, ctx)
        return hit


class GpuModel:
    """Owns an itjl_model (device copies of the model constants)."""

    def __init__(self, model, ctx):
        import ctypes
        self.ctx = ctx
        self.model = model
        d = _lib.ModelDesc()
        d.kind = _lib.KIND_OU_OSC if model.kind.startswith("one_timescale_and_osc") else _lib.KIND_OU
        if model.summary_method == "psd":
            # one_timescale_model: comp_psd (DSP periodogram); the oscillation models: comp_psd_adfriendly
            d.summary = _lib.SUMMARY_PGRAM if model.kind == "one_timescale" else _lib.SUMMARY_PSD
        elif model.kind.endswith("with_missing"):
            d.summary = _lib.SUMMARY_ACF_MISSING
        else:
            d.summary = _lib.SUMMARY_ACF
        d.distance = _lib.DIST_LINEAR if model.distance_method == "linear" else _lib.DIST_LOG
        d.update = _lib.UPDATE_EXACT if model.update == "exact" else _lib.UPDATE_EM
        d.n_trials = int(model.numTrials)
        d.n_t = int(model.n_t)
        d.dt = float(model.dt)
        # generate_data of the plain OU models never adds the mean (one_timescale.jl:226-228)
        d.data_mean = float(model.data_mean) if model.kind.startswith("one_timescale_and_osc") else 0.0
        d.data_sd = float(model.data_sd)
        self._target = np.ascontiguousarray(model.data_sum_stats, dtype=np.float64)
        d.n_stats = self._target.shape[0]
        d.target_stats = self._target.ctypes.data
        self._mask = self._bins = None
        if model.missing_mask is not None:
            self._mask = np.asfortranarray(np.asarray(model.missing_mask).astype(np.uint8))
            d.missing_mask = self._mask.ctypes.data
        if model.summary_method == "psd":
            self._bins = np.ascontiguousarray(np.nonzero(model.freq_idx)[0], dtype=np.int32)
            d.freq_bins = self._bins.ctypes.data
        self.n_stats = int(d.n_stats)
        self._h = ctypes.c_void_p()
        ctx.check(_lib.lib().itjl_model_create(ctx.handle, ctypes.byref(d), ctypes.byref(self._h)))
        ctx._models.add(self)

    @property
    def handle(self):
        return self._h

    def work(self):
        """(algorithmic flops per OU sample, samples per particle)."""
        import ctypes
        f = ctypes.c_double(0.0)
        s = ctypes.c_int64(0)
        self.ctx.check(_lib.lib().itjl_model_work(self._h, ctypes.byref(f), ctypes.byref(s)))
        return f.value, s.value

    def kernel_info(self):
        """(name of the fused kernel this model launches, tensor-core flops executed per OU sample)."""
        import ctypes
        name = ctypes.create_string_buffer(32)
        f = ctypes.c_double(0.0)
        self.ctx.check(_lib.lib().itjl_model_kernel_info(self._h, name, ctypes.byref(f)))
        return name.value.decode(), f.value

    def distances(self, theta, seed=0, generation=0, particle_offset=0, u0=None, noise=None,
                  phases=None, return_stats=False):
        """d[i] = generate_data_and_reduce(model, theta[i, :]) for a batch."""
        theta = np.atleast_2d(np.asarray(theta, dtype=np.float64))
        n, p = theta.shape
        th = np.asfortranarray(theta)
        out = np.empty(n, dtype=np.float64)
        stats = np.empty((n, self.n_stats), dtype=np.float64) if return_stats else None
        cast = lambda a: None if a is None else np.ascontiguousarray(a, dtype=np.float64)
        u0, noise, phases = cast(u0), cast(noise), cast(phases)
        combined = bool(self.model.distance_combined)
        if combined and stats is None and u0 is None and noise is None and phases is None:
            # the whole combined distance on the device (itjl_abc_distances_combined): the statistics never leave it
            m = self.model
            w = np.ascontiguousarray(m.weights, dtype=np.float64)
            lf = np.ascontiguousarray(m.lags_freqs, dtype=np.float64) if m.summary_method == "psd" else None
            self.ctx.check(_lib.lib().itjl_abc_distances_combined(
                self._h, _lib.ptr(th), n, p, int(seed), int(generation), int(particle_offset), _lib.ptr(lf), _lib.ptr(w),
                int(w.size), float(m.data_tau), float(m.data_osc) if m.data_osc is not None else 0.0, _lib.ptr(out)))
            return out
        if combined and stats is None:
            stats = np.empty((n, self.n_stats), dtype=np.float64)
        self.ctx.check(_lib.lib().itjl_abc_distances(
            self._h, _lib.ptr(th), n, p, int(seed), int(generation), int(particle_offset),
            _lib.ptr(u0), _lib.ptr(noise), _lib.ptr(phases), _lib.ptr(out), _lib.ptr(stats)))
        if combined:
            out = self._combine(out, stats)
        return (out, stats) if return_stats else out

    def _combine(self, d_stats, stats):
        """combined_distance: w1 * distance(stats) + w2 * (data_tau - tau_sim)^2 [+ w3 * (data_osc - osc_sim)^2].
        ACF (one_timescale.jl:278-290, :311-313): tau_sim = fit_expdecay(lags, simulated mean ACF); PSD
        (one_timescale.jl:314-316, one_timescale_and_osc.jl:377-384): tau_sim = tau_from_knee(knee) of the Lorentzian
        fit, osc_sim = find_oscillation_peak of its residual within freqlims - both fitted on the device for the
        whole batch (itjl_fit_expdecay / itjl_fit_knee)."""
        from . import utils
        m = self.model
        ok = np.isfinite(stats).all(axis=1)
        tau = np.full(stats.shape[0], np.nan)
        osc = np.full(stats.shape[0], np.nan)
        if ok.any():
            if m.summary_method == "acf":
                tau[ok] = np.atleast_1d(utils.fit_expdecay(m.lags_freqs, stats[ok], ctx=self.ctx))
            elif m.kind.startswith("one_timescale_and_osc"):
                # the knee is fitted with find_knee_frequency's defaults; the peak search window [freqlims] contains every
                # selected frequency (freq_idx is freqlims' open interval), i.e. it is the whole axis
                fit, peak = utils.find_knee_frequency(stats[ok], m.lags_freqs, ctx=self.ctx, return_peak=True)
                tau[ok] = utils.tau_from_knee(fit[:, 1])
                osc[ok] = peak
            else:
                tau[ok] = utils.tau_from_knee(utils.find_knee_frequency(stats[ok], m.lags_freqs, ctx=self.ctx)[:, 1])
        d = m.weights[0] * d_stats + m.weights[1] * (m.data_tau - tau) ** 2
        if m.summary_method == "psd" and m.kind.startswith("one_timescale_and_osc"):
            d = d + m.weights[2] * (m.data_osc - osc) ** 2
        return d

    def generation(self, theta, epsilon, min_accepted, seed=0, generation=0, particle_offset=0):
        """Distances + accept mask + sequential stop for a batch (abc.jl:181-199)."""
        import ctypes
        theta = np.atleast_2d(np.asarray(theta, dtype=np.float64))
        n, p = theta.shape
        th = np.asfortranarray(theta)
        isacc = np.empty(n, dtype=np.uint8)
        n_total = ctypes.c_int64(0)
        n_acc = ctypes.c_int64(0)
        if self.model.distance_combined:
            # distances need the per-particle exp-decay fit: score, combine, then the device accept / stop rule
            dist = np.ascontiguousarray(self.distances(theta, seed=seed, generation=generation,
                                                       particle_offset=particle_offset))
            self.ctx.check(_lib.lib().itjl_abc_accept(self.ctx.handle, _lib.ptr(dist), n, float(epsilon),
                                                      int(min_accepted), _lib.ptr(isacc), ctypes.byref(n_total),
                                                      ctypes.byref(n_acc)))
            return dist, isacc.astype(bool), int(n_total.value), int(n_acc.value)
        dist = np.empty(n, dtype=np.float64)
        self.ctx.check(_lib.lib().itjl_abc_generation(
            self._h, _lib.ptr(th), n, p, int(seed), int(generation), int(particle_offset),
            float(epsilon), int(min_accepted), _lib.ptr(dist), _lib.ptr(isacc),
            ctypes.byref(n_total), ctypes.byref(n_acc)))
        return dist, isacc.astype(bool), int(n_total.value), int(n_acc.value)

    def distances_dev(self, theta_dev_ptr, n_particles, n_theta, dist_dev_ptr, seed=0, generation=0,
                      particle_offset=0):
        """Device-pointer variant (asynchronous on the context stream)."""
        import ctypes
        self.ctx.check(_lib.lib().itjl_abc_distances_dev(
            self._h, ctypes.c_void_p(theta_dev_ptr), int(n_particles), int(n_theta), int(seed),
            int(generation), int(particle_offset), ctypes.c_void_p(dist_dev_ptr)))

    def distances_sharded_dev(self, theta_dev_ptr, n_particles, n_theta, dist_dev_ptr, seed=0, generation=0,
                              particle_offset=0):
        """Collective device-pointer variant: this rank scores its share of the n_particles in theta_dev, the
        library all-gathers, dist_dev receives all of them (itjl_abc_distances_sharded_dev)."""
        import ctypes
        self.ctx.check(_lib.lib().itjl_abc_distances_sharded_dev(
            self._h, ctypes.c_void_p(theta_dev_ptr), int(n_particles), int(n_theta), int(seed),
            int(generation), int(particle_offset), ctypes.c_void_p(dist_dev_ptr)))

    def close(self):
        if self._h:
            _lib.lib().itjl_model_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


# ---------------------------------------------------------- constructors ----
def check_model_inputs(data, time, fit_method, summary_method, prior, distance_method):
    """src/core/model.jl:239-292."""
    data = np.asarray(data, dtype=np.float64)
    if data.ndim == 1:
        data = data.reshape(1, -1)
    time = np.asarray(time, dtype=np.float64)
    if prior is not None and not isinstance(prior, (list, tuple)) and prior != "informed_prior":
        prior = [prior]
    if time.shape[0] != data.shape[-1]:
        raise ValueError("Time vector length must match data length")
    if np.any(np.diff(time) < 0):
        raise ValueError("Time vector must be monotonically increasing")
    if fit_method not in ("abc", "advi", "acw"):
        raise ValueError("fit_method must be :abc, :advi, or :acw")
    if fit_method == "advi":
        raise NotImplementedError("ADVI is outside the B200 hot path (SURVEY.md section 2)")
    if summary_method not in ("acf", "psd"):
        raise ValueError("summary_method must be :acf or :psd")
    if distance_method is not None and distance_method not in ("linear", "logarithmic"):
        raise ValueError("distance_method must be :linear or :logarithmic")
    return data, time, (None if prior in (None, "informed_prior") else list(prior))


def _acf_branch(acf_mean, n, dt, n_lags, prior):
    """one_timescale.jl:151-165."""
    if n_lags is None:
        a0 = utils.acw0(np.arange(n, dtype=np.float64), acf_mean)
        if math.isnan(a0):
            raise ValueError("InexactError: Int64(NaN) (the data ACF never crosses zero)")
        n_lags = int(math.floor(a0 * 1.1))
    lags = (np.arange(n, dtype=np.float64) * dt)[:n_lags]
    stats = np.ascontiguousarray(acf_mean[:n_lags])
    if prior is None:
        prior = [Normal(utils.fit_expdecay(lags, stats), 20.0)]          # informed_prior, :20-23
    return n_lags, lags, stats, prior


def _combined(model, distance_combined, weights, data_tau, ctx):
    """distance_combined = true of the ACF branch (one_timescale.jl:166-169, 278-290): the timescale of
    the data's mean ACF is fitted once at construction."""
    if not distance_combined:
        return model
    from . import utils
    w = [0.5, 0.5] if weights is None else [float(x) for x in weights]
    if len(w) != 2:
        raise ValueError("weights must have length 2 for the ACF summary")
    model.distance_combined = True
    model.weights = w
    model.data_tau = float(data_tau) if data_tau is not None else float(
        utils.fit_expdecay(model.lags_freqs, model.data_sum_stats, ctx=ctx))
    return model


def _one_timescale_psd(data, time, fit_method, prior, distance_method, freqlims, distance_combined, weights,
                       data_tau, update, dt, T, ctx):
    """PSD branch of one_timescale_model (one_timescale.jl:177-203): comp_psd = DSP periodogram, frequencies
    inside freqlims (default (0.5, 100)), logarithmic distance; informed prior and combined distance through the
    Lorentzian knee (tau_from_knee(find_knee_frequency(...)[2]))."""
    psd, freqs = summary.comp_psd(data, 1.0 / dt, ctx=ctx)
    mean_psd = psd.mean(axis=0)
    if freqlims is None:
        freqlims = (0.5, 100.0)
    freq_idx = (freqs < freqlims[1]) & (freqs > freqlims[0])
    lags_freqs = freqs[freq_idx]
    stats = np.ascontiguousarray(mean_psd[freq_idx])
    if prior is None:
        prior = [Normal(utils.tau_from_knee(utils.find_knee_frequency(stats, lags_freqs, ctx=ctx)[1]), 20.0)]
    m = TimescaleModel("one_timescale", data, time, fit_method, "psd", lags_freqs, prior,
                       distance_method or "logarithmic", stats, dt, T, data.shape[0],
                       float(data.mean()), float(data.std(ddof=1)), freqlims=freqlims, freq_idx=freq_idx,
                       update=update)
    if distance_combined:
        w = [0.5, 0.5] if weights is None else [float(x) for x in weights]
        if len(w) != 2:
            raise ValueError("weights must have length 2")
        m.distance_combined = True
        m.weights = w
        m.data_tau = float(data_tau) if data_tau is not None else float(
            utils.tau_from_knee(utils.find_knee_frequency(stats, lags_freqs, ctx=ctx)[1]))
    return m


def one_timescale_model(data, time, fit_method="abc", summary_method="acf", prior=None,
                        n_lags=None, distance_method=None, distance_combined=False,
                        weights=None, data_tau=None, update="exact", ctx=None, freqlims=None):
    data, time, prior = check_model_inputs(data, time, fit_method, summary_method, prior, distance_method)
    dt = float(time[1] - time[0]); T = float(time[-1])
    if summary_method == "psd":
        return _one_timescale_psd(data, time, fit_method, prior, distance_method, freqlims, distance_combined,
                                  weights, data_tau, update, dt, T, ctx)
    acf_mean = summary.comp_ac_fft(data).mean(axis=0)
    n_lags, lags, stats, prior = _acf_branch(acf_mean, data.shape[1], dt, n_lags, prior)
    m = TimescaleModel("one_timescale", data, time, fit_method, "acf", lags, prior,
                       distance_method or "linear", stats, dt, T, data.shape[0],
                       float(data.mean()), float(data.std(ddof=1)), n_lags=n_lags, update=update)
    return _combined(m, distance_combined, weights, data_tau, ctx)


def one_timescale_with_missing_model(data, time, fit_method="abc", summary_method="acf", prior=None,
                                     n_lags=None, distance_method=None, distance_combined=False,
                                     weights=None, data_tau=None, update="exact", ctx=None):
    data, time, prior = check_model_inputs(data, time, fit_method, summary_method, prior, distance_method)
    if summary_method != "acf":
        raise NotImplementedError("the Lomb-Scargle PSD branch is outside the B200 hot path")
    dt = float(time[1] - time[0]); T = float(time[-1])
    mask = np.isnan(data)
    acf_mean = summary.comp_ac_time_missing(data).mean(axis=0)
    n_lags, lags, stats, prior = _acf_branch(acf_mean, data.shape[1], dt, n_lags, prior)
    m = TimescaleModel("one_timescale_with_missing", data, time, fit_method, "acf", lags, prior,
                       distance_method or "linear", stats, dt, T, data.shape[0],
                       float(np.nanmean(data)), float(np.nanstd(data, ddof=1)), n_lags=n_lags,
                       missing_mask=mask, update=update)
    return _combined(m, distance_combined, weights, data_tau, ctx)


def _osc_informed_prior(stats, lags_freqs, summary_method, ctx):
    """informed_prior of the oscillation models (one_timescale_and_osc.jl:16-43): lorentzian_initial_guess (no fit)
    of the statistic - also for the ACF, as the reference does -, oscillation peak of the residual over the whole
    axis; PSD: Normal(1 / knee, 1), ACF: Normal(fit_expdecay, 20)."""
    amp, knee = utils.lorentzian_initial_guess(stats, lags_freqs, ctx=ctx)
    resid = np.asarray(stats, dtype=np.float64) - utils.lorentzian(lags_freqs, [amp, knee])
    osc_peak = utils.find_oscillation_peak(resid, lags_freqs, lags_freqs[0], lags_freqs[-1], ctx=ctx)
    if summary_method == "psd":
        return [Normal(1.0 / knee, 1.0), Normal(osc_peak, 0.1), Uniform(0.0, 1.0)]
    return [Normal(utils.fit_expdecay(lags_freqs, stats, ctx=ctx), 20.0), Normal(osc_peak, 0.1), Uniform(0.0, 1.0)]


def one_timescale_and_osc_model(data, time, fit_method="abc", summary_method="psd", prior=None,
                                n_lags=None, distance_method=None, freqlims=None,
                                distance_combined=False, weights=None, data_tau=None, data_osc=None,
                                update="exact", ctx=None):
    data, time, prior = check_model_inputs(data, time, fit_method, summary_method, prior, distance_method)
    dt = float(time[1] - time[0]); T = float(time[-1])
    mean, sd = float(data.mean()), float(data.std(ddof=1))
    w = [0.5, 0.5] if weights is None else [float(x) for x in weights]
    if summary_method == "acf":
        acf_mean = summary.comp_ac_fft(data).mean(axis=0)
        n_lags, lags, stats, _ = _acf_branch(acf_mean, data.shape[1], dt, n_lags, prior if prior is not None else [None])
        if prior is None:
            prior = _osc_informed_prior(stats, lags, "acf", ctx)
        m = TimescaleModel("one_timescale_and_osc", data, time, fit_method, "acf", lags, prior,
                           distance_method or "linear", stats, dt, T, data.shape[0], mean, sd,
                           n_lags=n_lags, update=update)
        if len(w) != 2 and distance_combined:
            raise ValueError("weights must have length 2 for the ACF summary")
        return _combined(m, distance_combined, w, data_tau, ctx)
    psd, freqs = summary.comp_psd_adfriendly(data, 1.0 / dt, ctx=ctx)
    mean_psd = psd.mean(axis=0)
    if freqlims is None:
        freqlims = (0.5, 100.0)
    freq_idx = (freqs < freqlims[1]) & (freqs > freqlims[0])
    lags_freqs = freqs[freq_idx]
    stats = np.ascontiguousarray(mean_psd[freq_idx])
    if prior is None:
        prior = _osc_informed_prior(stats, lags_freqs, "psd", ctx)
    m = TimescaleModel("one_timescale_and_osc", data, time, fit_method, "psd", lags_freqs, prior,
                       distance_method or "logarithmic", stats, dt, T, data.shape[0], mean, sd,
                       freqlims=freqlims, freq_idx=freq_idx, update=update)
    if distance_combined:
        # one_timescale_and_osc.jl:229-236; the PSD combination has three terms (:337-343)
        if len(w) == 2:
            raise ValueError("weights must have length 3 for the PSD summary of the oscillation model "
                             "(the reference's default [0.5, 0.5] throws a BoundsError at weights[3])")
        if len(w) != 3:
            raise ValueError("weights must have length 3 for the PSD summary")
        (amp, knee), _ = utils.find_knee_frequency(stats, lags_freqs, ctx=ctx, return_peak=True)
        resid = stats - utils.lorentzian(lags_freqs, [amp, knee])
        m.distance_combined = True
        m.weights = w
        m.data_tau = float(data_tau) if data_tau is not None else float(utils.tau_from_knee(knee))
        m.data_osc = float(data_osc) if data_osc is not None else float(
            utils.find_oscillation_peak(resid, lags_freqs, freqlims[0], freqlims[1], ctx=ctx))
    return m


def one_timescale_and_osc_with_missing_model(data, time, fit_method="abc", summary_method="acf", prior=None,
                                             n_lags=None, distance_method=None, distance_combined=False,
                                             update="exact"):
    """src/models/one_timescale_and_osc_with_missing.jl:159-222, ACF branch (the reference's default
    summary_method = :psd is the Lomb-Scargle branch, outside the B200 hot path)."""
    if distance_combined:
        raise NotImplementedError("distance_combined is outside the B200 hot path")
    data, time, prior = check_model_inputs(data, time, fit_method, summary_method, prior, distance_method)
    if summary_method != "acf":
        raise NotImplementedError("the Lomb-Scargle PSD branch is outside the B200 hot path")
    dt = float(time[1] - time[0]); T = float(time[-1])
    mask = np.isnan(data)
    acf_mean = summary.comp_ac_time_missing(data).mean(axis=0)
    n_lags, lags, stats, _ = _acf_branch(acf_mean, data.shape[1], dt, n_lags, prior if prior is not None else [None])
    if prior is None:
        # informed_prior, ACF branch (one_timescale_and_osc_with_missing.jl:36-48): the oscillation prior is the fixed
        # Normal(0.01, 0.05) there - the peak it computes is not used
        prior = [Normal(utils.fit_expdecay(lags, stats), 20.0), Normal(0.01, 0.05), Uniform(0.0, 1.0)]
    return TimescaleModel("one_timescale_and_osc_with_missing", data, time, fit_method, "acf", lags, prior,
                          distance_method or "linear", stats, dt, T, data.shape[0],
                          float(np.nanmean(data)), float(np.nanstd(data, ddof=1)), n_lags=n_lags,
                          missing_mask=mask, update=update)


# ---------------------------------------------------------------- protocol ----
def draw_theta(model, rng):
    """src/core/model.jl:154-156."""
    return np.array([p.rand(rng) for p in model.prior], dtype=np.float64)


def generate_data(model, theta, seed=0, ctx=None):
    """generate_data(model, theta): one synthetic data set, (numTrials, n_t)."""
    from . import ou
    if model.kind.startswith("one_timescale_and_osc"):
        data = ou.generate_ou_with_oscillation(theta, model.dt, model.T, model.numTrials,
                                               model.data_mean, model.data_sd, seed=seed,
                                               update=model.update, ctx=ctx)
        if model.kind.endswith("with_missing"):
            data[model.missing_mask] = np.nan
        return data
    data = ou.generate_ou_process(theta[0], model.data_sd, model.dt, model.T, model.numTrials,
                                  seed=seed, update=model.update, ctx=ctx)
    if model.kind == "one_timescale_with_missing":
        data[model.missing_mask] = np.nan
    return data


def summary_stats(model, data, ctx=None):
    if model.summary_method == "acf":
        if model.kind.endswith("with_missing"):
            return summary.comp_ac_time_missing(data, model.n_lags, ctx=ctx).mean(axis=0)
        return summary.comp_ac_fft(data, model.n_lags, ctx=ctx).mean(axis=0)
    if model.kind == "one_timescale":
        psd, _ = summary.comp_psd(data, 1.0 / model.dt, ctx=ctx)                 # one_timescale.jl:255
    else:
        psd, _ = summary.comp_psd_adfriendly(data, 1.0 / model.dt, ctx=ctx)
    return psd.mean(axis=0)[model.freq_idx]


def distance_function(model, sum_stats, data_sum_stats):
    """linear_distance / logarithmic_distance (src/stats/distances.jl:29-53) on host vectors
    (O(n_stats); the batched path computes them inside the fused kernel)."""
    a = np.asarray(sum_stats, dtype=np.float64)
    b = np.asarray(data_sum_stats, dtype=np.float64)
    if model.distance_method == "linear":
        return float(np.mean((a - b) ** 2))
    if model.distance_method == "logarithmic":
        with np.errstate(invalid="ignore", divide="ignore"):
            return float(np.mean((np.log(a) - np.log(b)) ** 2))
    raise ValueError("Distance method must be :linear or :logarithmic")


def generate_data_and_reduce(model, theta, seed=0, generation=0, particle=0, ctx=None):
    """src/core/model.jl:146-151 for ONE particle (a batch-of-one fused kernel launch)."""
    return float(model.gpu_model(ctx).distances(np.asarray(theta, dtype=np.float64).reshape(1, -1),
                                                seed=seed, generation=generation,
                                                particle_offset=particle)[0])
