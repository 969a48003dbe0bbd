"""Oracle: the ABC / PMC-ABC sampler.

Follows /root/reference/src/core/abc.jl:
  draw_theta_pmc           :101-117
  basic_abc                :156-215
  pmc_abc                  :287-502
  calc_weights             :520-567   (the multi-parameter branch :547-566 is the
                                       intended formula; the 1-parameter branch
                                       :533-544 indexes axes(theta_prev, 2) and an
                                       uninitialised buffer - restated here as the
                                       same Gaussian-mixture denominator)
  weighted_covar           :582-613
  effective_sample_size    :632-642
  select_epsilon           :700-761   (StatsBase.percentile = type-7 quantile)
  compute_adaptive_alpha   :807-837
  get_param_dict_abc       :932-974
`find_MAP` (KernelDensity) stays on the Julia side and is not restated.

`distance_fn(theta, particle_index, generation)` abstracts the per-particle
generate_data_and_reduce so tests can plug in the NumPy model, the C restatement or
pre-computed distances.
"""
import math

import numpy as np


def get_param_dict_abc():
    """abc.jl:932-974."""
    return dict(epsilon_0=1.0, max_iter=10000, min_accepted=100, steps=30,
                sample_only=False, minAccRate=0.01, target_acc_rate=0.01,
                target_epsilon=1e-4, show_progress=True, verbose=True, jitter=1e-6,
                distance_max=10.0, quantile_lower=25.0, quantile_upper=75.0,
                quantile_init=50.0, acc_rate_buffer=0.1, alpha_max=0.9, alpha_min=0.1,
                acc_rate_far=2.0, acc_rate_close=0.2, alpha_far_mult=1.5,
                alpha_close_mult=0.5, convergence_window=5, theta_rtol=1e-2,
                theta_atol=1e-3, N=10000)


def draw_theta_pmc(rng, theta_prev, weights, tau_squared, jitter=1e-5):
    """abc.jl:101-117 (jitter default 1e-5 is the function's own, not pmc_abc's)."""
    w = np.asarray(weights, dtype=np.float64)
    idx = rng.choice(theta_prev.shape[0], p=w / w.sum())
    theta_star = theta_prev[idx]
    cov = np.asarray(tau_squared, dtype=np.float64) + jitter * np.eye(theta_prev.shape[1])
    while True:
        theta = rng.multivariate_normal(theta_star, cov)
        if not np.any(theta < 0):
            return theta


def basic_abc(distance_fn, draw_fn, n_theta, epsilon, max_iter, min_accepted,
              generation=0, proposals=None, distance_batch_fn=None, lookahead=64):
    """abc.jl:156-215.  `draw_fn()` -> theta (prior or PMC proposal); if
    `proposals` (max_iter, n_theta) is given they are consumed in order instead.
    `distance_batch_fn(thetas, first_index, generation)` (optional) scores `lookahead`
    proposals at a time - proposals are i.i.d., so drawing a few ahead of the serial loop and
    discarding the unused tail leaves every returned quantity distributed as in the
    reference; it only lets the C restatement use all host cores."""
    samples = np.zeros((max_iter, n_theta))
    dist = np.zeros(max_iter)
    accepted = 0
    it = 0
    ahead_theta = ahead_d = None
    ahead_base = 0
    for i in range(max_iter):
        if distance_batch_fn is not None:
            if ahead_theta is None or i - ahead_base >= ahead_theta.shape[0]:
                nb = min(lookahead, max_iter - i)
                ahead_theta = (np.asarray(proposals[i:i + nb]) if proposals is not None
                               else np.stack([draw_fn() for _ in range(nb)]))
                ahead_d = distance_batch_fn(ahead_theta, i, generation)
                ahead_base = i
            theta = ahead_theta[i - ahead_base]
            d = ahead_d[i - ahead_base]
        else:
            theta = proposals[i] if proposals is not None else draw_fn()
            d = distance_fn(theta, i, generation)
        samples[i] = theta
        dist[i] = d
        it += 1
        if d <= epsilon:                                   # NaN -> False (:185)
            accepted += 1
        if accepted == min_accepted:                       # :190-194
            break
    isacc = dist[:it] <= epsilon
    return dict(samples=samples, isaccepted=isacc, theta_accepted=samples[:it][isacc],
                distances=dist[:it].copy(), n_accepted=int(isacc.sum()), n_total=it,
                epsilon=epsilon)


def stop_index(distances, epsilon, min_accepted):
    """n_total of basic_abc for a pre-computed distance vector: 1 + index of the
    min_accepted-th acceptance, else len(distances)."""
    acc = np.asarray(distances) <= epsilon
    c = np.cumsum(acc)
    hit = np.nonzero(c == min_accepted)[0]
    return int(hit[0]) + 1 if hit.size else len(acc)


def percentile_type7(x, p):
    """StatsBase.percentile(x, p) (quantile definition 7)."""
    xs = np.sort(np.asarray(x, dtype=np.float64))
    h = (xs.size - 1) * (p / 100.0)
    lo = int(math.floor(h))
    hi = min(lo + 1, xs.size - 1)
    return float(xs[lo] + (h - lo) * (xs[hi] - xs[lo]))


def compute_adaptive_alpha(iteration, current_acc_rate, target_acc_rate, alpha_max=0.9,
                           alpha_min=0.1, total_iterations=100, acc_rate_far=2.0,
                           acc_rate_close=0.2, alpha_far_mult=1.5, alpha_close_mult=0.5):
    """abc.jl:807-837."""
    progress = iteration / total_iterations
    base = alpha_max * (1 - progress) + alpha_min * progress
    diff = abs(current_acc_rate - target_acc_rate) / target_acc_rate
    if diff > acc_rate_far:
        return min(alpha_max, base * alpha_far_mult)
    if diff < acc_rate_close:
        return max(alpha_min, base * alpha_close_mult)
    return base


def select_epsilon(distances, current_epsilon, target_acc_rate=0.01, current_acc_rate=0.0,
                   iteration=1, total_iterations=100, distance_max=10.0,
                   quantile_lower=25.0, quantile_upper=75.0, quantile_init=50.0,
                   acc_rate_buffer=0.1, alpha_max=0.9, alpha_min=0.1, acc_rate_far=2.0,
                   acc_rate_close=0.2, alpha_far_mult=1.5, alpha_close_mult=0.5):
    """abc.jl:700-761.  When iteration > 1 and current_acc_rate == 0 the reference
    leaves `new_epsilon` undefined (UndefVarError, :746-760); the restatement keeps
    the current epsilon."""
    d = np.asarray(distances, dtype=np.float64)
    d = d[~np.isnan(d)]
    d = d[d < distance_max]
    if d.size == 0:
        return current_epsilon
    q_lo = percentile_type7(d, quantile_lower)
    q_init = percentile_type7(d, quantile_init)
    q_hi = percentile_type7(d, quantile_upper)
    alpha = compute_adaptive_alpha(iteration, current_acc_rate, target_acc_rate,
                                   alpha_max=alpha_max, alpha_min=alpha_min,
                                   total_iterations=total_iterations,
                                   acc_rate_far=acc_rate_far, acc_rate_close=acc_rate_close,
                                   alpha_far_mult=alpha_far_mult,
                                   alpha_close_mult=alpha_close_mult)
    if iteration == 1:
        return q_init
    new_eps = current_epsilon
    if current_acc_rate > 0.0:
        if current_acc_rate > target_acc_rate * (1.0 + acc_rate_buffer):
            new_eps = max(q_lo, current_epsilon * (1.0 - alpha))
        elif current_acc_rate < target_acc_rate * (1.0 - acc_rate_buffer):
            new_eps = min(q_hi, current_epsilon * (1.0 + alpha))
    return new_eps


def mvnormal_pdf(x, mean, cov):
    k = mean.shape[-1]
    diff = x - mean
    sol = np.linalg.solve(cov, diff.T).T
    quad = np.sum(diff * sol, axis=-1)
    return np.exp(-0.5 * quad) / math.sqrt((2 * math.pi) ** k * np.linalg.det(cov))


def calc_weights(theta_prev, theta, tau_squared, weights, prior):
    """abc.jl:520-567, intended formula:
    w_i ~ prod_j prior_j(theta_ij) / sum_j w_j N(theta_i; theta_prev_j, tau^2)."""
    theta_prev = np.atleast_2d(theta_prev)
    theta = np.atleast_2d(theta)
    cov = np.atleast_2d(np.asarray(tau_squared, dtype=np.float64))
    w = np.asarray(weights, dtype=np.float64)
    new = np.empty(theta.shape[0])
    for i in range(theta.shape[0]):
        p = 1.0
        for j, pr in enumerate(prior):
            p *= float(pr.pdf(theta[i, j]))
        kern = mvnormal_pdf(theta[i][None, :], theta_prev, cov)
        new[i] = p / np.sum(w * kern)
    return new / new.sum()


def weighted_covar(x, w):
    """abc.jl:582-613."""
    x = np.atleast_2d(x)
    wn = np.asarray(w, dtype=np.float64) / np.sum(w)
    sumw = wn.sum()
    if not math.isclose(sumw, 1.0, rel_tol=1e-10):
        return np.zeros((x.shape[1], x.shape[1]))
    sum2 = np.sum(wn ** 2)
    xbar = (wn[:, None] * x).sum(axis=0)
    xc = x - xbar
    cov = (xc * wn[:, None]).T @ xc
    return cov * sumw / (sumw * sumw - sum2)


def effective_sample_size(w):
    """abc.jl:632-642."""
    w = np.asarray(w, dtype=np.float64)
    return float(w.sum() ** 2 / np.sum(w ** 2))


def pmc_abc(model_prior, distance_fn, rng, n_theta=None, distance_batch_fn=None, epsilon_0=1.0, max_iter=10000,
            min_accepted=100, steps=10, sample_only=False, minAccRate=0.01,
            target_acc_rate=0.01, target_epsilon=5e-3, jitter=1e-6, distance_max=10.0,
            quantile_lower=25.0, quantile_upper=75.0, quantile_init=50.0,
            acc_rate_buffer=0.1, alpha_max=0.9, alpha_min=0.1, acc_rate_far=2.0,
            acc_rate_close=0.2, alpha_far_mult=1.5, alpha_close_mult=0.5,
            convergence_window=3, theta_rtol=1e-2, theta_atol=1e-3, **_ignored):
    """abc.jl:287-502.  Returns the list of per-step records (output_record)."""
    prior = list(model_prior)
    ndim = len(prior)
    eps_kw = dict(target_acc_rate=target_acc_rate, distance_max=distance_max,
                  quantile_lower=quantile_lower, quantile_upper=quantile_upper,
                  quantile_init=quantile_init, acc_rate_buffer=acc_rate_buffer,
                  alpha_max=alpha_max, alpha_min=alpha_min, acc_rate_far=acc_rate_far,
                  acc_rate_close=acc_rate_close, alpha_far_mult=alpha_far_mult,
                  alpha_close_mult=alpha_close_mult)
    records = []
    epsilon = epsilon_0
    theta_history = []
    for i_step in range(1, steps + 1):
        if i_step == 1:
            draw = lambda: np.array([p.rand(rng) for p in prior])
            res = basic_abc(distance_fn, draw, ndim, epsilon, max_iter, min_accepted,
                            generation=i_step, distance_batch_fn=distance_batch_fn)
            epsilon = select_epsilon(res["distances"], epsilon, **eps_kw)
            theta = res["theta_accepted"]
            tau_squared = 2.0 * np.atleast_2d(np.cov(theta, rowvar=False))
            tau_squared = tau_squared + jitter * np.eye(ndim)
            weights = np.full(theta.shape[0], 1.0 / theta.shape[0])
            eff = effective_sample_size(weights)
        else:
            prev = records[-1]
            draw = lambda: draw_theta_pmc(rng, prev["theta_accepted"], prev["weights"],
                                          prev["tau_squared"])
            res = basic_abc(distance_fn, draw, ndim, epsilon, max_iter, min_accepted,
                            generation=i_step, distance_batch_fn=distance_batch_fn)
            theta = res["theta_accepted"]
            eff = effective_sample_size(prev["weights"])
            if sample_only:
                weights = np.zeros(0); tau_squared = np.zeros((0, 0))
            else:
                weights = calc_weights(prev["theta_accepted"], theta, prev["tau_squared"],
                                       prev["weights"], prior)
                tau_squared = 2.0 * weighted_covar(theta, weights)
            acc_rate = res["n_accepted"] / res["n_total"]
            epsilon = select_epsilon(res["distances"], epsilon, current_acc_rate=acc_rate,
                                     iteration=i_step, total_iterations=steps, **eps_kw)
        records.append(dict(theta_accepted=theta, D_accepted=res["distances"],
                            n_accepted=res["n_accepted"], n_total=res["n_total"],
                            epsilon=epsilon, weights=weights, tau_squared=tau_squared,
                            eff_sample=eff))
        current_theta = theta.mean(axis=0, keepdims=True)
        accept_rate = res["n_accepted"] / res["n_total"]
        if accept_rate < minAccRate:                               # :461
            return records
        theta_history.append(current_theta)
        if len(theta_history) >= convergence_window:               # :469-491
            converged = True
            for dim in range(ndim):
                recent = [float(h[:, dim].mean()) for h in theta_history[-convergence_window:]]
                rng_ = max(recent) - min(recent)
                mean_ = float(np.mean(recent))
                if not (rng_ / abs(mean_) < theta_rtol or rng_ < theta_atol):
                    converged = False
                    break
            if converged:
                return records
        if epsilon < target_epsilon:                               # :494
            return records
    return records
