"""ABC / PMC-ABC driver with the per-particle loop body on the GPU.

Mirrors src/core/abc.jl (names, keyword arguments, defaults, return fields):
  basic_abc               :156-215   serial rejection loop  ->  batched: proposals are drawn
                                     for a chunk, the fused kernel scores the chunk, K3 applies
                                     `d <= epsilon` and the sequential stopping rule, and the
                                     chunk is truncated at the min_accepted-th acceptance, so
                                     n_total / acceptance rates are distributed as in the serial code
  draw_theta_pmc          :101-117   vectorised (same resample + MvNormal + redraw-while-negative)
  pmc_abc                 :287-502
  calc_weights            :520-567   (Gaussian-mixture denominator; the reference's 1-parameter
                                      branch :533-544 reads an uninitialised buffer - see DESIGN.md)
  weighted_covar          :582-613
  effective_sample_size   :632-642
  select_epsilon          :700-761
  compute_adaptive_alpha  :807-837
  find_MAP                :852-873   (Gaussian KDE on a random grid; KernelDensity.jl's
                                      bandwidth rule is not restated - not on the hot path)
  get_param_dict_abc      :932-974
  ABCResults/abc_results  :34-84
"""
import math
from dataclasses import dataclass

import numpy as np

from . import _lib


def get_param_dict_abc():
    return dict(epsilon_0=1.0, max_iter=10000, min_accepted=100, steps=30, sample_only=False,
                minAccRate=0.01, target_acc_rate=0.01, target_epsilon=1e-4, show_progress=True,
                verbose=True, jitter=1e-6, distance_max=10.0, quantile_lower=25.0,
                quantile_upper=75.0, quantile_init=50.0, acc_rate_buffer=0.1, alpha_max=0.9,
                alpha_min=0.1, acc_rate_far=2.0, acc_rate_close=0.2, alpha_far_mult=1.5,
                alpha_close_mult=0.5, convergence_window=5, theta_rtol=1e-2, theta_atol=1e-3,
                N=10000)


@dataclass
class ABCResults:
    theta_history: list
    epsilon_history: list
    acc_rate_history: list
    weights_history: list
    final_theta: np.ndarray
    final_weights: np.ndarray
    MAP: np.ndarray


# ------------------------------------------------------------- proposals ----
def draw_theta_prior(model, rng, n):
    """n i.i.d. draws of [rand(p) for p in prior] (model.jl:154-156) -> (n, n_theta)."""
    return np.stack([np.asarray(p.rand(rng, n), dtype=np.float64) for p in model.prior], axis=1)


def draw_theta_pmc(model, theta_prev, weights, tau_squared, rng, n=1, jitter=1e-5):
    """n i.i.d. PMC proposals (abc.jl:101-117): theta* resampled with the weights, then
    MvNormal(theta*, tau_squared + jitter I), redrawn while any component is negative."""
    theta_prev = np.atleast_2d(np.asarray(theta_prev, dtype=np.float64))
    w = np.asarray(weights, dtype=np.float64)
    p = theta_prev.shape[1]
    cov = np.atleast_2d(np.asarray(tau_squared, dtype=np.float64)) + jitter * np.eye(p)
    chol = np.linalg.cholesky(cov)
    if n >= 4096:
        # large batches (the 10^6-particle sweep): multinomial counts + a random order are the same i.i.d.
        # resampling as n inverse-cdf look-ups, without n cache-missing binary searches
        idx = rng.permutation(np.repeat(np.arange(w.size), rng.multinomial(n, w / w.sum())))
    else:
        cdf = np.cumsum(w / w.sum())
        idx = np.minimum(np.searchsorted(cdf, rng.random(n), side="right"), w.size - 1)
    stars = theta_prev[idx]
    out = stars + rng.standard_normal((n, p)) @ chol.T
    bad = np.nonzero(np.any(out < 0, axis=1))[0]
    while bad.size:                                   # redraw around the SAME theta*
        out[bad] = stars[bad] + rng.standard_normal((bad.size, p)) @ chol.T
        bad = bad[np.any(out[bad] < 0, axis=1)]
    return out


def draw_theta_pmc_device(theta_prev, weights, tau_squared, n, seed=0, generation=0, first=0, jitter=1e-5, ctx=None):
    """The same n i.i.d. PMC proposals drawn on the device (itjl_pmc_propose: inverse-cdf resampling + Gaussian
    perturbation + redraw-while-negative on the Philox stream (seed, generation, first + i)).  For large generations:
    10^6 proposals cost ~1 ms instead of the 60-90 ms of the NumPy path, and every rank of a multi-GPU run draws the
    same matrix without a host RNG in lock step."""
    theta_prev = np.asfortranarray(np.atleast_2d(np.asarray(theta_prev, dtype=np.float64)))
    w = np.ascontiguousarray(weights, dtype=np.float64)
    n_prev, p = theta_prev.shape
    tsq = np.asfortranarray(np.atleast_2d(np.asarray(tau_squared, dtype=np.float64)))
    out = np.empty((int(n), p), dtype=np.float64, order="F")
    ctx = ctx or _lib.default_context()
    ctx.check(_lib.lib().itjl_pmc_propose(ctx.handle, _lib.ptr(theta_prev), n_prev, p, _lib.ptr(w), _lib.ptr(tsq), float(jitter),
                                          int(n), int(seed), int(generation), int(first), _lib.ptr(out)))
    return out


# ------------------------------------------------------------- basic_abc ----
def _first_chunk(max_iter, min_accepted, sm_count):
    return int(min(max_iter, max(2 * sm_count, 2 * min_accepted)))


def basic_abc(model, epsilon, max_iter, min_accepted, pmc_mode=False, weights=None, theta_prev=None,
              tau_squared=None, show_progress=False, rng=None, seed=0, generation=0, ctx=None,
              scorer=None, proposals=None, device_proposals=False):
    """Returns the reference's NamedTuple as a dict: samples, isaccepted, theta_accepted,
    distances, n_accepted, n_total, epsilon, weights, tau_squared, eff_sample.

    `scorer(theta_chunk, epsilon, needed, generation, particle_offset)` ->
    (dist, isaccepted, n_total, n_accepted) defaults to the single-GPU fused kernels; the
    multi-GPU sharded scorer lives in distributed.py.  `proposals` (max_iter, n_theta)
    overrides the proposal draws (parity tests); `device_proposals` draws the PMC proposals with
    draw_theta_pmc_device instead of the host RNG (same distribution, Philox stream of `seed`)."""
    rng = rng or np.random.default_rng()
    n_theta = len(model.prior)
    if scorer is None:
        gm = model.gpu_model(ctx)
        sm = gm.ctx.sm_count

        def scorer(th, eps, needed, gen, off):
            return gm.generation(th, eps, needed, seed=seed, generation=gen, particle_offset=off)
    else:
        sm = 148
    samples = np.zeros((max_iter, n_theta))
    distances = np.zeros(max_iter)
    offset = 0
    accepted = 0
    it = 0
    chunk = _first_chunk(max_iter, min_accepted, sm)
    while offset < max_iter:
        n = min(chunk, max_iter - offset)
        if proposals is not None:
            th = np.asarray(proposals[offset:offset + n], dtype=np.float64)
        elif pmc_mode and device_proposals:
            th = draw_theta_pmc_device(theta_prev, weights, tau_squared, n, seed=seed, generation=generation, first=offset, ctx=ctx)
        elif pmc_mode:
            th = draw_theta_pmc(model, theta_prev, weights, tau_squared, rng, n)
        else:
            th = draw_theta_prior(model, rng, n)
        needed = min_accepted - accepted
        d, _, n_tot, n_acc = scorer(th, float(epsilon), int(needed), int(generation), int(offset))
        samples[offset:offset + n_tot] = th[:n_tot]
        distances[offset:offset + n_tot] = d[:n_tot]
        accepted += n_acc
        it = offset + n_tot
        offset += n
        if needed > 0 and n_acc >= needed:
            break                                           # abc.jl:190-194
        rate = accepted / offset if offset else 0.0
        if rate > 0.0:
            want = int(1.25 * (min_accepted - accepted) / rate) + 1
            chunk = int(min(max(want, 2 * sm), 4 * chunk))
        else:
            chunk = 4 * chunk
    dist = distances[:it].copy()
    isacc = dist <= epsilon
    theta_accepted = samples[:it][isacc]
    return dict(samples=samples, isaccepted=isacc, theta_accepted=theta_accepted, distances=dist,
                n_accepted=int(isacc.sum()), n_total=int(it), epsilon=float(epsilon),
                weights=np.ones(theta_accepted.size), tau_squared=np.zeros((theta_accepted.size,) * 2)
                if theta_accepted.size < 4096 else None, eff_sample=theta_accepted.size)


# ------------------------------------------------------ PMC bookkeeping ----
def percentile(x, p):
    """StatsBase.percentile (quantile type 7)."""
    return float(np.percentile(np.asarray(x, dtype=np.float64), p, method="linear"))


def compute_adaptive_alpha(iteration, current_acc_rate, target_acc_rate, alpha_max=0.9, alpha_min=0.1,
                           total_iterations=100, acc_rate_far=2.0, acc_rate_close=0.2,
                           alpha_far_mult=1.5, alpha_close_mult=0.5):
    progress = iteration / total_iterations
    base_alpha = alpha_max * (1 - progress) + alpha_min * progress
    acc_rate_diff = abs(current_acc_rate - target_acc_rate) / target_acc_rate
    if acc_rate_diff > acc_rate_far:
        return min(alpha_max, base_alpha * alpha_far_mult)
    if acc_rate_diff < acc_rate_close:
        return max(alpha_min, base_alpha * alpha_close_mult)
    return base_alpha


def select_epsilon(distances, current_epsilon, target_acc_rate=0.01, current_acc_rate=0.0, iteration=1,
                   total_iterations=100, distance_max=10.0, quantile_lower=25.0, quantile_upper=75.0,
                   quantile_init=50.0, acc_rate_buffer=0.1, alpha_max=0.9, alpha_min=0.1,
                   acc_rate_far=2.0, acc_rate_close=0.2, alpha_far_mult=1.5, alpha_close_mult=0.5, ctx=None):
    d = np.ascontiguousarray(distances, dtype=np.float64)
    if ctx is not None and d.size >= 65536:
        # large generations: the order statistics come from a device sort (itjl_order_stats); same numbers - the
        # interpolation below is NumPy's own (quantile type 7 = StatsBase.percentile)
        import ctypes
        probs = np.array([quantile_lower, quantile_init, quantile_upper], dtype=np.float64) / 100.0
        pairs = np.empty(6, dtype=np.float64)
        nv = ctypes.c_int64(0)
        ctx.check(_lib.lib().itjl_order_stats(ctx.handle, _lib.ptr(d), d.size, float(distance_max), _lib.ptr(probs), 3,
                                              _lib.ptr(pairs), ctypes.byref(nv)))
        if nv.value == 0:
            return current_epsilon
        qs = []
        for i, p in enumerate(probs):
            pos = p * (nv.value - 1)
            t = pos - math.floor(pos)
            a, b = pairs[2 * i], pairs[2 * i + 1]
            v = a + (b - a) * t
            if t >= 0.5:                                  # numpy's _lerp
                v = b - (b - a) * (1.0 - t)
            qs.append(float(v))
        q_lower, q_init, q_upper = qs
    else:
        valid = d[~np.isnan(d)]
        valid = valid[valid < distance_max]
        if valid.size == 0:
            return current_epsilon
        # one selection pass for the three order statistics (StatsBase.percentile = quantile type 7)
        q_lower, q_init, q_upper = (float(q) for q in np.percentile(valid, [quantile_lower, quantile_init, quantile_upper],
                                                                   method="linear"))
    alpha = compute_adaptive_alpha(iteration, current_acc_rate, target_acc_rate, alpha_max=alpha_max,
                                   alpha_min=alpha_min, total_iterations=total_iterations,
                                   acc_rate_far=acc_rate_far, acc_rate_close=acc_rate_close,
                                   alpha_far_mult=alpha_far_mult, alpha_close_mult=alpha_close_mult)
    if iteration == 1:
        return q_init
    if current_acc_rate > 0.0:
        if current_acc_rate > target_acc_rate * (1.0 + acc_rate_buffer):
            return max(q_lower, current_epsilon * (1.0 - alpha))
        if current_acc_rate < target_acc_rate * (1.0 - acc_rate_buffer):
            return min(q_upper, current_epsilon * (1.0 + alpha))
    return current_epsilon            # reference: UndefVarError when the rate is exactly 0


def calc_weights(theta_prev, theta, tau_squared, weights, prior, ctx=None, normalize=True):
    """calc_weights (abc.jl:520-567): w_i ~ prod_k pdf(prior_k, theta_ik) / sum_j w_j N(theta_i; theta_prev_j, tau^2).
    The prior densities are evaluated here (the host owns the distribution objects, as Julia does); the
    n x n_prev Gaussian mixture sum runs on the device (itjl_pmc_weights, k_pmc_mixture) - on a multi-GPU
    context sharded over the rows and all-gathered."""
    import ctypes
    theta_prev = np.atleast_2d(np.asarray(theta_prev, dtype=np.float64))
    theta = np.atleast_2d(np.asarray(theta, dtype=np.float64))
    if theta_prev.shape[0] == 1 and theta_prev.shape[1] != len(prior):
        theta_prev = theta_prev.T
    if theta.shape[0] == 1 and theta.shape[1] != len(prior):
        theta = theta.T
    k = theta.shape[1]
    cov = np.asfortranarray(np.atleast_2d(np.asarray(tau_squared, dtype=np.float64)))
    w = np.ascontiguousarray(weights, dtype=np.float64)
    pri = np.ones(theta.shape[0])
    for j, p in enumerate(prior):
        pri *= p.pdf(theta[:, j])
    pri = np.ascontiguousarray(pri)
    tp, th = np.asfortranarray(theta_prev), np.asfortranarray(theta)
    out = np.empty(theta.shape[0], dtype=np.float64)
    ctx = ctx or _lib.default_context()
    ctx.check(_lib.lib().itjl_pmc_weights(ctx.handle, _lib.ptr(tp), tp.shape[0], _lib.ptr(th), th.shape[0], int(k),
                                          _lib.ptr(cov), _lib.ptr(w), _lib.ptr(pri), int(bool(normalize)), _lib.ptr(out)))
    return out


def weighted_covar(x, w, ctx=None):
    """weighted_covar (abc.jl:582-613) on the device (itjl_weighted_covar)."""
    x = np.asfortranarray(np.atleast_2d(np.asarray(x, dtype=np.float64)))
    w = np.ascontiguousarray(w, dtype=np.float64)
    out = np.empty((x.shape[1], x.shape[1]), dtype=np.float64, order="F")
    ctx = ctx or _lib.default_context()
    ctx.check(_lib.lib().itjl_weighted_covar(ctx.handle, _lib.ptr(x), x.shape[0], x.shape[1], _lib.ptr(w), _lib.ptr(out)))
    return np.ascontiguousarray(out)


def effective_sample_size(w):
    w = np.asarray(w, dtype=np.float64)
    return float(w.sum() ** 2 / np.sum(w ** 2))


def find_MAP(theta_accepted, N=10000, rng=None):
    from scipy.stats import gaussian_kde
    rng = rng or np.random.default_rng()
    theta_accepted = np.atleast_2d(np.asarray(theta_accepted, dtype=np.float64))
    out = np.empty(theta_accepted.shape[1])
    for i in range(theta_accepted.shape[1]):
        col = theta_accepted[:, i]
        pos = rng.uniform(col.min(), col.max(), N)
        if col.size < 2 or col.min() == col.max():
            out[i] = col[0]
            continue
        out[i] = pos[int(np.argmax(gaussian_kde(col, bw_method="silverman")(pos)))]
    return out


def abc_results(output_record, rng=None):
    final_theta = output_record[-1]["theta_accepted"]
    return ABCResults([r["theta_accepted"] for r in output_record],
                      [r["epsilon"] for r in output_record],
                      [r["n_accepted"] / r["n_total"] for r in output_record],
                      [r["weights"] for r in output_record],
                      final_theta, output_record[-1]["weights"],
                      find_MAP(final_theta, rng=rng) if final_theta.size else np.full(final_theta.shape[1], np.nan))


def pmc_abc(model, epsilon_0=1.0, max_iter=10000, min_accepted=100, steps=10, sample_only=False,
            minAccRate=0.01, target_acc_rate=0.01, target_epsilon=5e-3, show_progress=False,
            verbose=False, jitter=1e-6, distance_max=10.0, quantile_lower=25.0, quantile_upper=75.0,
            quantile_init=50.0, acc_rate_buffer=0.1, alpha_max=0.9, alpha_min=0.1, acc_rate_far=2.0,
            acc_rate_close=0.2, alpha_far_mult=1.5, alpha_close_mult=0.5, convergence_window=3,
            theta_rtol=1e-2, theta_atol=1e-3, rng=None, seed=0, ctx=None, scorer=None,
            return_record=False, calc_weights_fn=None, weighted_covar_fn=None, device_proposals=False, **_unused):
    """pmc_abc (abc.jl:287-502).  `scorer`, `calc_weights_fn(theta_prev, theta, tau_squared, weights, prior)` and
    `weighted_covar_fn(x, w)` are injection points for host-logic tests that run without a GPU; by default the
    particles are scored by the fused kernels and the weights / covariance by the device kernels of this
    module - there is no CPU implementation of either in the product."""
    rng = rng or np.random.default_rng(seed)
    cw = calc_weights_fn or (lambda a, b, c, d, e: calc_weights(a, b, c, d, e, ctx=ctx))
    wc = weighted_covar_fn or (lambda x, w: weighted_covar(x, w, ctx=ctx))
    eps_kw = dict(target_acc_rate=target_acc_rate, distance_max=distance_max,
                  quantile_lower=quantile_lower, quantile_upper=quantile_upper,
                  quantile_init=quantile_init, acc_rate_buffer=acc_rate_buffer, alpha_max=alpha_max,
                  alpha_min=alpha_min, acc_rate_far=acc_rate_far, acc_rate_close=acc_rate_close,
                  alpha_far_mult=alpha_far_mult, alpha_close_mult=alpha_close_mult)
    record = []
    epsilon = float(epsilon_0)
    theta_history = []
    ndim = len(model.prior)

    def finish():
        res = abc_results(record, rng=rng)
        return (res, record) if return_record else res

    for i_step in range(1, steps + 1):
        if verbose:
            print(f"Starting step {i_step}\nepsilon = {epsilon}")
        if i_step == 1:
            result = basic_abc(model, epsilon, max_iter, min_accepted, pmc_mode=False, rng=rng,
                               seed=seed, generation=i_step, ctx=ctx, scorer=scorer)
            epsilon = select_epsilon(result["distances"], epsilon, **eps_kw)
            theta = result["theta_accepted"]
            if theta.shape[0] >= 2:
                tau_squared = 2.0 * np.atleast_2d(np.cov(theta, rowvar=False)) + jitter * np.eye(ndim)
            else:
                # nothing (or one particle) accepted: the reference's cov() of that is NaN and its
                # fill(1 / 0, 0) an empty vector; it leaves through the minAccRate exit below
                tau_squared = np.full((ndim, ndim), np.nan)
            weights = np.full(theta.shape[0], 1.0 / theta.shape[0]) if theta.shape[0] else np.zeros(0)
            eff = effective_sample_size(weights) if theta.shape[0] else float("nan")
        else:
            prev = record[-1]
            result = basic_abc(model, epsilon, max_iter, min_accepted, pmc_mode=True,
                               weights=prev["weights"], theta_prev=prev["theta_accepted"],
                               tau_squared=prev["tau_squared"], rng=rng, seed=seed, generation=i_step,
                               ctx=ctx, scorer=scorer, device_proposals=device_proposals)
            theta = result["theta_accepted"]
            eff = effective_sample_size(prev["weights"])
            if sample_only:
                weights = np.zeros(0); tau_squared = np.zeros((0, 0))
            else:
                if theta.shape[0]:
                    weights = cw(prev["theta_accepted"], theta, prev["tau_squared"], prev["weights"], model.prior)
                    tau_squared = 2.0 * np.atleast_2d(wc(theta, weights))
                else:
                    weights = np.zeros(0); tau_squared = np.full((ndim, ndim), np.nan)
            current_acc_rate = result["n_accepted"] / result["n_total"]
            epsilon = select_epsilon(result["distances"], epsilon, current_acc_rate=current_acc_rate,
                                     iteration=i_step, total_iterations=steps, **eps_kw)
        record.append(dict(theta_accepted=theta, D_accepted=result["distances"],
                           n_accepted=result["n_accepted"], n_total=result["n_total"], epsilon=epsilon,
                           weights=weights, tau_squared=tau_squared, eff_sample=eff))
        current_theta = theta.mean(axis=0, keepdims=True) if theta.size else np.full((1, ndim), np.nan)
        accept_rate = result["n_accepted"] / result["n_total"]
        if verbose:
            print(f"Acceptance Rate = {accept_rate}\nCurrent theta = {current_theta}\n--------------------")
        if accept_rate < minAccRate:
            return finish()
        theta_history.append(current_theta)
        if len(theta_history) >= convergence_window:
            converged = True
            for dim in range(ndim):
                recent = [float(h[:, dim].mean()) for h in theta_history[-convergence_window:]]
                param_range = max(recent) - min(recent)
                param_mean = float(np.mean(recent))
                if not (param_range / abs(param_mean) < theta_rtol or param_range < theta_atol):
                    converged = False
                    break
            if converged:
                if verbose:
                    print(f"Parameters converged after {i_step} iterations")
                return finish()
        if epsilon < target_epsilon:
            return finish()
    return finish()


def int_fit(model, param_dict=None, **kw):
    """Models.int_fit(model, param_dict) for fit_method = :abc (one_timescale.jl:352-401)."""
    if model.fit_method != "abc":
        raise NotImplementedError("only fit_method=:abc runs on the B200 path")
    params = get_param_dict_abc()
    if param_dict:
        params.update(param_dict)
    params.update(kw)
    params.pop("N", None)
    return pmc_abc(model, **params)


def posterior_predictive(container, model, n_samples=100, rng=None, seed=0, ctx=None):
    """Numerical part of posterior_predictive (ext/IntPlottingExt/IntPlottingExt.jl:105-128): summary
    statistics of data simulated at `n_samples` rows drawn with replacement from the posterior, their mean and
    2.5 % / 97.5 % quantiles.  One batched launch of the fused kernel instead of n_samples generate_data +
    summary_stats calls; the plotting stays with the caller.  Returns a dict with `theta_samples`,
    `predictions` (n_samples, n_stats), `pred_mean`, `pred_lower`, `pred_upper`, `x` (= model.lags_freqs)."""
    rng = rng or np.random.default_rng()
    final_theta = np.atleast_2d(np.asarray(container.final_theta, dtype=np.float64))
    idx = rng.integers(0, final_theta.shape[0], size=int(n_samples))
    theta = final_theta[idx]
    plain = model
    if model.distance_combined:                 # the statistics do not depend on the distance
        import copy
        plain = copy.copy(model); plain.distance_combined = False; plain._gpu = {}
    _, stats = plain.gpu_model(ctx).distances(theta, seed=seed, generation=0, return_stats=True)
    return {"theta_samples": theta, "predictions": stats, "pred_mean": stats.mean(axis=0),
            "pred_lower": np.array([percentile(stats[:, i], 2.5) for i in range(stats.shape[1])]),
            "pred_upper": np.array([percentile(stats[:, i], 97.5) for i in range(stats.shape[1])]),
            "x": np.asarray(model.lags_freqs)}
