"""Oracle: the model protocol for the three in-scope models.

Follows
  /root/reference/src/core/model.jl:146-156 (generate_data_and_reduce, draw_theta),
  /root/reference/src/models/one_timescale.jl:135-175 (ACF branch of the
      constructor), :226-228, :251-253, :311-327,
  /root/reference/src/models/one_timescale_with_missing.jl:137-186, :244-248,
      :270-272, :333-349,
  /root/reference/src/models/one_timescale_and_osc.jl:98-189, :263-266, :289-297,
      :368-393.
The PSD branch of one_timescale_model is the DSP periodogram (summary.comp_psd_periodogram); informed priors
and combined distances that need the Lorentzian knee go through oracle/knee.py.
Out of scope: the Lomb-Scargle PSD branch of the missing-data MODELS (the stored-data
periodogram is in summary.py).
"""
import math
from dataclasses import dataclass, field

import numpy as np

from . import acwred, distances, ou, summary


# ------------------------------------------------------------------ priors ----
@dataclass
class Normal:
    mu: float
    sigma: float

    def rand(self, rng, size=None):
        return rng.normal(self.mu, self.sigma, size)

    def pdf(self, x):
        z = (np.asarray(x, dtype=np.float64) - self.mu) / self.sigma
        return np.exp(-0.5 * z * z) / (self.sigma * math.sqrt(2.0 * math.pi))


@dataclass
class Uniform:
    a: float
    b: float

    def rand(self, rng, size=None):
        return rng.uniform(self.a, self.b, size)

    def pdf(self, x):
        x = np.asarray(x, dtype=np.float64)
        return np.where((x >= self.a) & (x <= self.b), 1.0 / (self.b - self.a), 0.0)


# ------------------------------------------------------------------- models ----
@dataclass
class Model:
    kind: str                 # "one_timescale" | "one_timescale_with_missing" | "one_timescale_and_osc"
    data: np.ndarray
    time: np.ndarray
    summary_method: str       # "acf" | "psd"
    distance_method: str      # "linear" | "logarithmic"
    prior: list
    lags_freqs: np.ndarray
    data_sum_stats: np.ndarray
    dt: float
    T: float
    numTrials: int
    data_mean: float
    data_sd: float
    n_lags: int = None
    freqlims: tuple = None
    freq_idx: np.ndarray = None
    missing_mask: np.ndarray = None
    update: str = "exact"
    distance_combined: bool = False
    weights: list = None
    data_tau: float = None
    data_osc: float = None
    n_t: int = field(default=None)

    def __post_init__(self):
        if self.n_t is None:
            self.n_t = ou.n_time_points(self.dt, self.T)


def _check_inputs(data, time, prior):
    """model.jl:239-292 (the parts that apply to dims = last)."""
    data = np.asarray(data, dtype=np.float64)
    if data.ndim == 1:
        data = data.reshape(1, -1)
    time = np.asarray(time, dtype=np.float64)
    if time.shape[0] != data.shape[-1]:
        raise ValueError("Time vector length must match data length")
    if np.any(np.diff(time) < 0):
        raise ValueError("Time vector must be monotonically increasing")
    if prior is not None and not isinstance(prior, (list, tuple)):
        prior = [prior]
    return data, time, (list(prior) if prior is not None else None)


def _acf_setup(acf_mean, n, dt, n_lags):
    """one_timescale.jl:151-157."""
    if n_lags is None:
        lags_samples = np.arange(n, dtype=np.float64)
        a0 = acwred.acw0(lags_samples, acf_mean)
        if math.isnan(a0):
            raise ValueError("InexactError: Int64(NaN)")
        n_lags = int(math.floor(a0 * 1.1))
    lags = (np.arange(n, dtype=np.float64) * dt)[:n_lags]
    return n_lags, lags, acf_mean[:n_lags]


def with_combined_distance(model, weights=(0.5, 0.5), data_tau=None, data_osc=None):
    """distance_combined = true.  ACF (one_timescale.jl:166-169): data_tau = fit_expdecay(lags, data_sum_stats); PSD of
    one_timescale_model (:195-198): data_tau = tau_from_knee(find_knee_frequency(...)[2]); PSD of the oscillation model
    (one_timescale_and_osc.jl:229-236): that and data_osc = find_oscillation_peak of the Lorentzian's residual within
    freqlims, three weights."""
    from . import knee
    model.distance_combined = True
    model.weights = [float(w) for w in weights]
    if model.summary_method == "acf":
        model.data_tau = float(data_tau) if data_tau is not None else acwred.fit_expdecay(model.lags_freqs, model.data_sum_stats)
        return model
    u = knee.find_knee_frequency(model.data_sum_stats, model.lags_freqs)
    model.data_tau = float(data_tau) if data_tau is not None else float(knee.tau_from_knee(u[1]))
    if model.kind.startswith("one_timescale_and_osc"):
        resid = model.data_sum_stats - knee.lorentzian(model.lags_freqs, u)
        model.data_osc = float(data_osc) if data_osc is not None else knee.find_oscillation_peak(
            resid, model.lags_freqs, model.freqlims[0], model.freqlims[1])
    return model


def osc_informed_prior(stats, lags_freqs, summary_method):
    """informed_prior of the oscillation models (one_timescale_and_osc.jl:16-43)."""
    from . import knee
    u0 = knee.lorentzian_initial_guess(stats, lags_freqs)
    resid = np.asarray(stats) - knee.lorentzian(lags_freqs, u0)
    peak = knee.find_oscillation_peak(resid, lags_freqs, lags_freqs[0], lags_freqs[-1])
    if summary_method == "psd":
        return [Normal(1.0 / u0[1], 1.0), Normal(peak, 0.1), Uniform(0.0, 1.0)]
    return [Normal(acwred.fit_expdecay(lags_freqs, stats), 20.0), Normal(peak, 0.1), Uniform(0.0, 1.0)]


def one_timescale_model(data, time, fit_method="abc", summary_method="acf", prior=None,
                        n_lags=None, distance_method=None, update="exact", freqlims=None):
    """one_timescale.jl:135-175 (ACF branch), :177-203 (PSD branch: comp_psd = DSP periodogram)."""
    data, time, prior = _check_inputs(data, time, prior)
    dt = float(time[1] - time[0]); T = float(time[-1])
    if summary_method == "psd":
        psd, freqs = summary.comp_psd_periodogram(data, 1.0 / dt)              # :179-180
        mean_psd = psd.mean(axis=0)
        if freqlims is None:
            freqlims = (0.5, 100.0)                                           # :181-183
        freq_idx = (freqs < freqlims[1]) & (freqs > freqlims[0])              # :184
        lags_freqs, stats = freqs[freq_idx], mean_psd[freq_idx]
        if prior is None:                                                     # :187-190 -> :24-26
            from . import knee
            prior = [Normal(knee.tau_from_knee(knee.find_knee_frequency(stats, lags_freqs)[1]), 20.0)]
        return Model("one_timescale", data, time, "psd", distance_method or "logarithmic", prior,
                     lags_freqs, stats, dt, T, data.shape[0], float(data.mean()), float(data.std(ddof=1)),
                     freqlims=freqlims, freq_idx=freq_idx, update=update)
    acf_mean = summary.comp_ac_fft(data).mean(axis=0)
    n_lags, lags, stats = _acf_setup(acf_mean, data.shape[1], dt, n_lags)
    if prior is None:
        prior = [Normal(acwred.fit_expdecay(lags, stats), 20.0)]      # :20-23
    return Model("one_timescale", data, time, "acf", distance_method or "linear", prior,
                 lags, stats, dt, T, data.shape[0], float(data.mean()),
                 float(data.std(ddof=1)), n_lags=n_lags, update=update)


def one_timescale_with_missing_model(data, time, fit_method="abc", summary_method="acf",
                                     prior=None, n_lags=None, distance_method=None,
                                     update="exact"):
    """one_timescale_with_missing.jl:137-186 (ACF branch)."""
    if summary_method != "acf":
        raise NotImplementedError("Lomb-Scargle PSD branch is out of scope")
    data, time, prior = _check_inputs(data, time, prior)
    dt = float(time[1] - time[0]); T = float(time[-1])
    mask = np.isnan(data)
    acf_mean = summary.comp_ac_time_missing(data).mean(axis=0)
    n_lags, lags, stats = _acf_setup(acf_mean, data.shape[1], dt, n_lags)
    if prior is None:
        prior = [Normal(acwred.fit_expdecay(lags, stats), 20.0)]
    return Model("one_timescale_with_missing", data, time, "acf",
                 distance_method or "linear", prior, lags, stats, dt, T, data.shape[0],
                 float(np.nanmean(data)), float(np.nanstd(data, ddof=1)),
                 n_lags=n_lags, missing_mask=mask, update=update)


def one_timescale_and_osc_model(data, time, fit_method="abc", summary_method="psd",
                                prior=None, n_lags=None, distance_method=None,
                                freqlims=None, update="exact"):
    """one_timescale_and_osc.jl:98-189."""
    data, time, prior = _check_inputs(data, time, prior)
    dt = float(time[1] - time[0]); T = float(time[-1])
    mean, sd = float(data.mean()), float(data.std(ddof=1))
    if summary_method == "acf":
        acf_mean = summary.comp_ac_fft(data).mean(axis=0)
        n_lags, lags, stats = _acf_setup(acf_mean, data.shape[1], dt, n_lags)
        if prior is None:
            prior = osc_informed_prior(stats, lags, "acf")
        return Model("one_timescale_and_osc", data, time, "acf",
                     distance_method or "linear", prior, lags, stats, dt, T,
                     data.shape[0], mean, sd, n_lags=n_lags, update=update)
    psd, freqs = summary.comp_psd_adfriendly(data, 1.0 / dt)
    mean_psd = psd.mean(axis=0)
    if freqlims is None:
        freqlims = (0.5, 100.0)
    freq_idx = (freqs < freqlims[1]) & (freqs > freqlims[0])
    if prior is None:
        prior = osc_informed_prior(mean_psd[freq_idx], freqs[freq_idx], "psd")
    return Model("one_timescale_and_osc", data, time, "psd",
                 distance_method or "logarithmic", prior, freqs[freq_idx],
                 mean_psd[freq_idx], dt, T, data.shape[0], mean, sd,
                 freqlims=freqlims, freq_idx=freq_idx, update=update)


def one_timescale_and_osc_with_missing_model(data, time, fit_method="abc", summary_method="acf",
                                             prior=None, n_lags=None, distance_method=None,
                                             update="exact"):
    """one_timescale_and_osc_with_missing.jl:159-222 (ACF branch; the PSD branch is Lomb-Scargle,
    out of scope)."""
    if summary_method != "acf":
        raise NotImplementedError("Lomb-Scargle PSD branch is out of scope")
    data, time, prior = _check_inputs(data, time, prior)
    if prior is None:
        raise NotImplementedError("informed prior needs the Lorentzian fit (out of scope)")
    dt = float(time[1] - time[0]); T = float(time[-1])
    mask = np.isnan(data)                                                      # :177
    acf_mean = summary.comp_ac_time_missing(data).mean(axis=0)                 # :179-180
    n_lags, lags, stats = _acf_setup(acf_mean, data.shape[1], dt, n_lags)      # :183-187
    return Model("one_timescale_and_osc_with_missing", data, time, "acf",
                 distance_method or "linear", prior, lags, stats, dt, T, data.shape[0],
                 float(np.nanmean(data)), float(np.nanstd(data, ddof=1)),
                 n_lags=n_lags, missing_mask=mask, update=update)


# ------------------------------------------------------------------ protocol ----
def draw_theta(model, rng):
    """model.jl:154-156."""
    return np.array([p.rand(rng) for p in model.prior], dtype=np.float64)


def generate_data(model, theta, **rng_spec):
    """generate_data methods; rng_spec = dict(seed, generation, particle) or explicit
    u0/noise/phases."""
    if model.kind in ("one_timescale_and_osc", "one_timescale_and_osc_with_missing"):
        data = ou.generate_ou_with_oscillation(theta, model.dt, model.T, model.numTrials,
                                               model.data_mean, model.data_sd,
                                               update=model.update, n_t=model.n_t,
                                               **rng_spec)
        if model.kind == "one_timescale_and_osc_with_missing":
            data = data.copy()
            data[model.missing_mask] = np.nan             # and_osc_with_missing.jl:295
        return data
    rng_spec.pop("phases", None)
    data = ou.generate_ou_process(theta[0], model.data_sd, model.dt, model.T,
                                  model.numTrials, update=model.update, n_t=model.n_t,
                                  **rng_spec)
    if model.kind == "one_timescale_with_missing":
        data = data.copy()
        data[model.missing_mask] = np.nan                 # with_missing.jl:246
    return data


def summary_stats(model, data):
    if model.summary_method == "acf":
        if model.kind in ("one_timescale_with_missing", "one_timescale_and_osc_with_missing"):
            return summary.comp_ac_time_missing(data, model.n_lags).mean(axis=0)
        return summary.comp_ac_fft(data, model.n_lags).mean(axis=0)
    if model.kind == "one_timescale":
        psd, _ = summary.comp_psd_periodogram(data, 1.0 / model.dt)              # one_timescale.jl:255
    else:
        psd, _ = summary.comp_psd_adfriendly(data, 1.0 / model.dt)
    return psd.mean(axis=0)[model.freq_idx]


def distance_function(model, sum_stats, data_sum_stats):
    if model.distance_combined:
        # one_timescale.jl:306-320 + combined_distance :278-290; one_timescale_and_osc.jl:368-385 + :324-347
        from . import knee
        d1 = (distances.linear_distance(sum_stats, data_sum_stats) if model.distance_method == "linear"
              else distances.logarithmic_distance(sum_stats, data_sum_stats))
        if not np.all(np.isfinite(sum_stats)):
            return float("nan")
        if model.summary_method == "acf":
            tau_sim = acwred.fit_expdecay(model.lags_freqs, sum_stats)
            return model.weights[0] * d1 + model.weights[1] * distances.linear_distance(model.data_tau, tau_sim)
        u = knee.find_knee_frequency(sum_stats, model.lags_freqs)
        d = model.weights[0] * d1 + model.weights[1] * distances.linear_distance(model.data_tau, float(knee.tau_from_knee(u[1])))
        if model.kind.startswith("one_timescale_and_osc"):
            resid = np.asarray(sum_stats) - knee.lorentzian(model.lags_freqs, u)
            osc = knee.find_oscillation_peak(resid, model.lags_freqs, model.freqlims[0], model.freqlims[1])
            d += model.weights[2] * distances.linear_distance(model.data_osc, osc)
        return d
    if model.distance_method == "linear":
        return distances.linear_distance(sum_stats, data_sum_stats)
    if model.distance_method == "logarithmic":
        return distances.logarithmic_distance(sum_stats, data_sum_stats)
    raise ValueError("Distance method must be :linear or :logarithmic")


def generate_data_and_reduce(model, theta, **rng_spec):
    """model.jl:146-151."""
    synth = generate_data(model, theta, **rng_spec)
    with np.errstate(invalid="ignore", divide="ignore"):
        stats = summary_stats(model, synth)
        return distance_function(model, stats, model.data_sum_stats)
