"""Oracle: ACW reducers and fits.

Follows /root/reference/src/utils/utils.jl:
  expdecay / residual_expdecay!  :53-67
  fit_expdecay                   :113-138
  tau_from_acw50                 :196
  acw50 / acw0 / acweuler        :218-226, :257-264, :294-302
  acw_romberg                    :332-345

Third-party arithmetic that is NOT under /root/reference (restated from the
published algorithm; parity unpinned beyond the reference's own KATs):
  * Romberg.jl 0.2 (+ Richardson.jl 1.x, Primes.jl): `romberg(dx, y)` builds
    trapezoid sums at strides given by the prime factorisation of length(y)-1 and
    Richardson-extrapolates them (power = 2), returning the tableau entry with the
    smallest error estimate.  KATs: test/test_utils.jl:335-349.
  * NonlinearSolve 4 `LevenbergMarquardt()` on the scalar residual
    r(tau) = mean((exp(-lags/tau) - acf)^2) (utils.jl:63-67, :121-125).  LM on a
    scalar residual minimises r^2, whose stationary points are those of r; the oracle
    returns the argmin of r (bracketed regula falsi on dr/dlog(tau)), not LM's last
    iterate.  KATs: test/test_utils.jl:119-160 (rtol 0.01 / 0.1).
"""
import math

import numpy as np

INV_E = 1.0 / math.e          # Julia `1 / ℯ` = 0.36787944117144233


def first_crossing_index(acf, thr):
    """0-based index of the first acf[k] <= thr, or -1 (the reference returns
    lags[findfirst(acf .<= thr)] or NaN with a warning)."""
    acf = np.asarray(acf, dtype=np.float64)
    hit = np.nonzero(acf <= thr)[0]
    return int(hit[0]) if hit.size else -1


def _lag_or_nan(lags, k):
    return float(lags[k]) if k >= 0 else float("nan")


def acw0(lags, acf):
    """utils.jl:257-264."""
    return _lag_or_nan(lags, first_crossing_index(acf, 0.0))


def acw50(lags, acf):
    """utils.jl:218-226."""
    return _lag_or_nan(lags, first_crossing_index(acf, 0.5))


def acweuler(lags, acf):
    """utils.jl:294-302."""
    return _lag_or_nan(lags, first_crossing_index(acf, INV_E))


def tau_from_acw50(a50):
    """utils.jl:196: -acw50 / log(0.5)."""
    return -a50 / math.log(0.5)


# ---------------------------------------------------------------- Romberg ----
def prime_factors(m):
    """Primes.factor(m) flattened in increasing order (with multiplicity)."""
    out = []
    p = 2
    while p * p <= m:
        while m % p == 0:
            out.append(p)
            m //= p
        p += 1
    if m > 1:
        out.append(m)
    return out


def romberg_strides(n):
    """Strides of the trapezoid sums Romberg.jl forms for length(y) == n:
    1, f1, f1*f2, ..., n-1."""
    steps = [1]
    for f in prime_factors(n - 1):
        steps.append(steps[-1] * f)
    return steps


def richardson_extrapolate(vals, hs, power=2, breaktol=2.0,
                           rtol=math.sqrt(np.finfo(np.float64).eps), atol=0.0):
    """Richardson.jl `extrapolate(fh::Vector{Tuple}; power)`: Neville tableau over
    (value, h) pairs ordered by decreasing h; returns (estimate, error)."""
    f0 = vals[0]
    err = math.inf
    nev = [vals[0]]
    hp = [hs[0] ** power]
    for j in range(1, len(vals)):
        nev.append(vals[j])
        hp.append(hs[j] ** power)
        minerr = math.inf
        for i in range(len(nev) - 2, -1, -1):
            old = nev[i]
            nev[i] = nev[i + 1] + (nev[i + 1] - nev[i]) / (hp[i] / hp[-1] - 1.0)
            e = abs(nev[i] - old)
            minerr = min(minerr, e)
            if e < err:
                f0, err = nev[i], e
        if minerr > breaktol * err or not math.isfinite(minerr):
            break
        if err <= max(rtol * abs(f0), atol):
            break
    return f0, err


def romberg(dx, y):
    """Romberg.jl 0.2 `romberg(dx, y)` -> (integral, error)."""
    y = np.asarray(y, dtype=np.float64)
    n = y.shape[0]
    if n <= 1:
        raise ValueError("y array must have length > 1")
    endsum = 0.5 * (y[0] + y[-1])
    vals, hs = [], []
    for step in romberg_strides(n):
        inner = y[step:n - 1:step].sum() if step < n - 1 else 0.0
        vals.append(step * dx * (inner + endsum))
        hs.append(step * dx)
    vals.reverse()
    hs.reverse()
    return richardson_extrapolate(vals, hs, power=2)


def acw_romberg(dt, acf):
    """utils.jl:332-334."""
    return romberg(dt, acf)[0]


# ------------------------------------------------------------ exp-decay fit ----
def expdecay(tau, lags):
    """utils.jl:53-55."""
    return np.exp(-(1.0 / tau) * np.asarray(lags, dtype=np.float64))


def expdecay_residual(tau, lags, acf):
    """utils.jl:63-67: mean(abs2.(expdecay(tau, lags) .- acf))."""
    return float(np.mean((expdecay(tau, lags) - np.asarray(acf)) ** 2))


FIT_LOG_HALF_WIDTH = 12.0     # search bracket: log(tau0) +- this
FIT_BRACKET_STEP = 0.5
FIT_MAX_ITER = 100
FIT_TOL = 1e-13


def _dres_dlogtau(u, lags, acf):
    """d/du of sum((exp(-t e^{-u}) - y)^2), u = log tau (common factor 2 dropped)."""
    s = math.exp(-u)
    e = np.exp(-s * lags)
    return float(np.sum((e - acf) * e * lags) * s)


def fit_expdecay(lags, acf):
    """utils.jl:113-126 restated as the argmin of r(tau) over tau > 0.

    Algorithm (identical in the CUDA epilogue, csrc/acw_kernels.cu):
      u0 = log(acw50/ln2) (or 0 when no 50 % crossing, utils.jl:115-120);
      walk outwards from u0 in steps of 0.5 (<= 12 each way) until dr/du changes
      sign; Illinois regula falsi (bisection fallback) on dr/du inside that
      bracket until the iterate moves by < 1e-13 relative.  If no sign change
      exists within the walk, the last point walked to is returned.
    """
    lags = np.asarray(lags, dtype=np.float64)
    acf = np.asarray(acf, dtype=np.float64)
    a50 = acw50(lags, acf)
    tau0 = tau_from_acw50(a50)
    if not (tau0 > 0.0) or not math.isfinite(tau0):
        tau0 = 1.0
    u0 = math.log(tau0)
    g0 = _dres_dlogtau(u0, lags, acf)
    if g0 == 0.0:
        return math.exp(u0)
    # g > 0: residual increases with tau -> minimum lies below; walk down.
    direction = -1.0 if g0 > 0.0 else 1.0
    lo, glo = u0, g0
    hi, ghi = u0, g0
    found = False
    nsteps = int(FIT_LOG_HALF_WIDTH / FIT_BRACKET_STEP)
    u_prev, g_prev = u0, g0
    for _ in range(nsteps):
        u = u_prev + direction * FIT_BRACKET_STEP
        g = _dres_dlogtau(u, lags, acf)
        if g == 0.0:
            return math.exp(u)
        if (g > 0.0) != (g_prev > 0.0):
            if direction < 0:
                lo, glo, hi, ghi = u, g, u_prev, g_prev
            else:
                lo, glo, hi, ghi = u_prev, g_prev, u, g
            found = True
            break
        u_prev, g_prev = u, g
    if not found:
        return math.exp(u_prev)
    # bracket [lo, hi] with g(lo) < 0 < g(hi): regula-falsi/bisection hybrid
    # (Illinois), deterministic and derivative-free beyond g itself.
    a, ga, b, gb = lo, glo, hi, ghi
    side = 0
    u = 0.5 * (a + b)
    u_last = a
    for _ in range(FIT_MAX_ITER):
        u = (a * gb - b * ga) / (gb - ga)
        if not (a < u < b):
            u = 0.5 * (a + b)
        g = _dres_dlogtau(u, lags, acf)
        if g == 0.0 or abs(u - u_last) < FIT_TOL * max(1.0, abs(u)):
            break
        u_last = u
        if (g > 0.0) == (gb > 0.0):
            b, gb = u, g
            if side == -1:
                ga *= 0.5
            side = -1
        else:
            a, ga = u, g
            if side == 1:
                gb *= 0.5
            side = 1
    return math.exp(u)


def fit_expdecay_3_parameters(lags, acf):
    """utils.jl:159-179: A (exp(-t / tau) + B) fitted to acf[2:end] over lags[2:end] with the scalar residual mean(abs2(.)),
    from u0 = [0.5, tau_from_acw50(acw50(lags, acf)), 0] ([0.5, 1, 0] when the ACF never falls to 0.5); tau < 0 -> NaN.
    NonlinearSolve's LM is not restated: this returns the argmin of the same objective (for a fixed tau the best
    (A, A B) is a closed-form linear least-squares pair; the remaining 1-D problem in log tau is minimised over
    tau0 / 32 .. 32 tau0 by a 65-point log scan + golden section).  PARITY UNPINNED."""
    from scipy import optimize
    lags = np.asarray(lags, dtype=np.float64); acf = np.asarray(acf, dtype=np.float64)
    a50 = acw50(lags, acf)
    tau0 = 1.0 if math.isnan(a50) else tau_from_acw50(a50)
    t, y = lags[1:], acf[1:]
    if t.size < 3 or not np.all(np.isfinite(y)) or not tau0 > 0:
        return float("nan")

    def reduced(lt):
        e = np.exp(-t * math.exp(-lt))
        A = np.stack([e, np.ones_like(e)], axis=1)
        coef, *_ = np.linalg.lstsq(A, y, rcond=None)
        r = A @ coef - y
        return float(np.mean(r * r))
    # the objective in log tau can have more than one dip: scan tau0 / 32 .. 32 tau0 on 65 log-spaced points, then
    # golden-section search between the neighbours of the best one (the same procedure as csrc/knee_kernels.cu)
    x0, span = math.log(tau0), math.log(32.0)
    grid = [x0 - span + i * (span / 32.0) for i in range(65)]
    vals = [reduced(g) for g in grid]
    i = int(np.argmin(vals))
    lo, hi = x0 - span + (i - 1) * (span / 32.0), x0 - span + (i + 1) * (span / 32.0)
    gr = 0.6180339887498949
    c1, c2 = hi - gr * (hi - lo), lo + gr * (hi - lo)
    f1, f2 = reduced(c1), reduced(c2)
    for _ in range(70):
        if f1 < f2:
            hi, c2, f2 = c2, c1, f1
            c1 = hi - gr * (hi - lo); f1 = reduced(c1)
        else:
            lo, c1, f1 = c1, c2, f2
            c2 = lo + gr * (hi - lo); f2 = reduced(c2)
    return float(math.exp(0.5 * (lo + hi)))
