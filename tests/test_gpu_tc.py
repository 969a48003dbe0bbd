"""GPU parity of the tensor-core (tcgen05 / TMEM) fused ACF kernel: the operand-layout probe, and
k_abc_acf_tc against the fp64 oracle and against the CUDA-core kernel k_abc_acf on the same noise."""
import ctypes
import os

import numpy as np
import pytest

from oracle import models as omodels
from oracle import ou as oou

pytestmark = pytest.mark.gpu

DT = 0.002
ACF_TOL = 1e-5        # north_star (1): ACF within 1e-5 (values in [-1, 1])
SBO = 256             # probe images: one K step (16 rows x 16 B) per 8-wide MN block
REGION = 16 * SBO


def idesc(m, n):
    return (1 << 4) | (1 << 15) | (1 << 16) | ((n >> 3) << 17) | ((m >> 4) << 24)


def probe(pkg, ctx, img, ops, lbo, sbo, idsc):
    out = np.zeros((128, 512), dtype=np.float32)
    ops = np.ascontiguousarray(np.asarray(ops, dtype=np.int32).reshape(-1, 4))
    ctx.check(pkg._lib.lib().itjl_tc_probe(ctx.handle, img.ctypes.data_as(ctypes.c_void_p), img.nbytes,
                                           ops.ctypes.data_as(ctypes.c_void_p), len(ops), lbo, sbo, idsc,
                                           out.ctypes.data_as(ctypes.c_void_p)))
    return out


def test_probe_row_shifted_gram_is_exact(pkg, ctx):
    """MN-major no-swizzle operands, B = the same buffer started s K-rows later: integer Gram
    products X[:rows]^T X[s:s+rows] come back exactly (what the lag sums are built on)."""
    T, ksteps, n_shift = 128, 2, 4
    rows = 16 * ksteps
    rows_total = rows + n_shift - 1
    rng = np.random.default_rng(0)
    x = rng.integers(-4, 5, size=rows_total * T).astype(np.float64)
    sbo = rows_total * 16
    img = np.zeros(16 * sbo // 2, dtype=np.float16)
    j = np.arange(x.size)
    i, a = j // T, j % T
    img[((a // 8) * sbo + i * 16 + (a % 8) * 2) // 2] = x
    X = x.reshape(rows_total, T)
    ops = [[k * 256, (k * 16 + s) * 16, 128 * s, 1 if k > 0 else 0] for k in range(ksteps) for s in range(n_shift)]
    out = probe(pkg, ctx, img, ops, 128, sbo, idesc(128, 128))
    for s in range(n_shift):
        assert np.array_equal(out[:, 128 * s:128 * (s + 1)], X[:rows].T @ X[s:s + rows])


def test_probe_m64_lane_mapping_and_lane_offset(pkg, ctx):
    """An M = 64 MMA writes row r to TMEM lane 32 (r / 16) + r % 16 (+ the D address' lane offset);
    tensor-memory plan B of k_abc_acf_tc puts rows a >= 64 of D_0 and D_4 on the upper 16 lanes."""
    img = np.zeros(2 * REGION // 2, dtype=np.float16)
    a = np.arange(128)
    img[((a // 8) * SBO + (a % 8) * 2) // 2] = a + 1                       # A[k=0][a] = a + 1
    img[(REGION + (a // 8) * SBO + (a % 8) * 2) // 2] = 1.0                # B[k=0][b] = 1
    out = probe(pkg, ctx, img, [[0, REGION, 0, 0], [8 * SBO, REGION, (16 << 16) | 0, 0]], 128, SBO, idesc(64, 32))
    r = np.arange(64)
    lanes = 32 * (r // 16) + r % 16
    assert np.array_equal(out[lanes, 0], r + 1.0)              # rows 0..63 on the lower half lanes
    assert np.array_equal(out[lanes + 16, 0], r + 65.0)        # A started 8 blocks later, lane offset 16
    assert np.array_equal(out[lanes, 31], r + 1.0)


def test_probe_accumulator_truncates_toward_zero(pkg, ctx):
    """Documents why k_abc_acf_tc drains its accumulators every segment: products are aligned to the
    fp32 accumulator and truncated toward zero (not rounded to nearest), also within one MMA."""
    img = np.zeros(4 * REGION // 2, dtype=np.float16)
    a = np.arange(128)
    pos = lambda region, k: (region * REGION + (a // 8) * SBO + k * 16 + (a % 8) * 2) // 2
    img[pos(2, 0)] = 4096.0
    img[pos(3, 0)] = np.where(a % 2 == 0, 4096.0, -4096.0)                  # D = +-2^24
    img[pos(0, 0)] = 1.75
    img[pos(1, 0)] = np.where(a % 2 == 0, 1.0, -1.0)
    out = probe(pkg, ctx, img, [[2 * REGION, 3 * REGION, 0, 0], [0, REGION, 0, 1]], 128, SBO, idesc(128, 32))
    assert out[0, 0] == 2.0 ** 24 and out[0, 1] == -2.0 ** 24             # RN would give +-(2^24 + 2)
    for k in range(16):
        img[pos(0, k)] = 0.25
        img[pos(1, k)] = 1.0
    out = probe(pkg, ctx, img, [[2 * REGION, 3 * REGION, 0, 0], [0, REGION, 0, 1]], 128, SBO, idesc(128, 32))
    assert out[0, 0] == 2.0 ** 24                                           # 16 x 0.25 = 4 is lost term by term


def _models(pkg, n_t, n_trials, n_lags, missing=False, seed=3):
    data = oou.generate_ou_process(0.3, 1.0, DT, n_t * DT, n_trials, seed=seed, n_t=n_t)
    t = DT * np.arange(1, n_t + 1)
    if missing:
        rng = np.random.default_rng(seed)
        for r in range(n_trials):
            s = rng.integers(0, n_t - n_t // 10)
            data[r, s:s + n_t // 10] = np.nan
        om = omodels.one_timescale_with_missing_model(data, t, prior=[omodels.Uniform(0.05, 5)], n_lags=n_lags)
        mk = lambda: pkg.one_timescale_with_missing_model(data, t, "abc", prior=[pkg.Uniform(0.05, 5)], n_lags=n_lags)
    else:
        om = omodels.one_timescale_model(data, t, prior=[omodels.Uniform(0.05, 5)], n_lags=n_lags)
        mk = lambda: pkg.one_timescale_model(data, t, "abc", prior=[pkg.Uniform(0.05, 5)], n_lags=n_lags)
    return om, mk


def _gpu_models(mk, ctx, **tuning):
    """(tensor-core model, CUDA-core model): the kernel is chosen at itjl_model_create from the context's
    itjl_tuning (abc_tc = 1 / 0; further knobs in `tuning`)."""
    with ctx.tuned(abc_tc=1, **tuning):
        g_tc = mk().gpu_model(ctx)
    with ctx.tuned(abc_tc=0):
        g_ff = mk().gpu_model(ctx)
    assert g_tc.kernel_info()[0] == "k_abc_acf_tc" and g_tc.kernel_info()[1] > 0
    assert g_ff.kernel_info() == ("k_abc_acf", 0.0)
    return g_tc, g_ff


# (n_t, n_trials, n_lags): tensor-memory plan A with 1..4 tiles, plan B (M = 64 split, 5th tile), several
# accumulator segments (21 and 50 trials), fewer trials than groups, odd buffer counts, FFMA tail lags
# (one lag tile), two to four lag windows (n_lags 465 .. 1537: further launches with the B operand 3 rows down each)
SHAPES = [(1000, 5, 37), (601, 3, 20), (2500, 1, 300), (5000, 2, 385), (5000, 2, 386), (5000, 4, 414),
          (5000, 21, 449), (5000, 3, 430), (3000, 50, 200), (5000, 50, 414), (5000, 7, 401),
          (5000, 3, 463), (5000, 9, 470), (5000, 5, 509),
          (5000, 5, 510), (5000, 3, 600), (5000, 12, 769), (3000, 4, 700),
          (5000, 3, 1000), (5000, 2, 1537), (3000, 4, 1200)]


# tc_tail_max = 60: the two- and four-tile FFMA tails (no longer chosen by default: a second lag window is faster)
@pytest.mark.parametrize("n_t,n_trials,n_lags,tail_max", [s + (None,) for s in SHAPES] + [(5000, 9, 470, 60), (5000, 5, 509, 60)])
def test_tc_kernel_matches_oracle_and_cuda_core_kernel(pkg, ctx, n_t, n_trials, n_lags, tail_max):
    om, mk = _models(pkg, n_t, n_trials, n_lags)
    g_tc, g_ff = _gpu_models(mk, ctx, **({} if tail_max is None else {"tc_tail_max": tail_max}))
    rng = np.random.default_rng(n_t + n_lags)
    theta = np.array([[0.3], [4.9], [0.05], [1.2]])
    P = theta.shape[0]
    u0 = rng.standard_normal((P, n_trials))
    noise = rng.standard_normal((P, n_trials, n_t))
    d_tc, s_tc = g_tc.distances(theta, u0=u0, noise=noise, return_stats=True)
    d_ff, s_ff = g_ff.distances(theta, u0=u0, noise=noise, return_stats=True)
    for i in range(P):
        ac = omodels.summary_stats(om, omodels.generate_data(om, theta[i], u0=u0[i], noise=noise[i]))
        assert np.max(np.abs(s_tc[i] - ac)) <= ACF_TOL, (i, np.max(np.abs(s_tc[i] - ac)))
        d_ref = np.mean((ac - om.data_sum_stats) ** 2)
        assert abs(d_tc[i] - d_ref) <= 2 * np.sqrt(d_ref) * ACF_TOL + ACF_TOL ** 2 + 1e-5 * d_ref
    assert np.max(np.abs(s_tc - s_ff)) <= ACF_TOL
    # Philox mode: same stream in both kernels, same batching invariance
    th = rng.uniform(0.05, 5.0, size=(40, 1))
    a = g_tc.distances(th, seed=11, generation=2)
    b = g_ff.distances(th, seed=11, generation=2)
    assert np.all(np.abs(a - b) <= 2 * np.sqrt(np.maximum(b, 0)) * ACF_TOL + ACF_TOL ** 2 + 1e-5 * b)
    parts = np.concatenate([g_tc.distances(th[:13], seed=11, generation=2),
                            g_tc.distances(th[13:], seed=11, generation=2, particle_offset=13)])
    assert np.array_equal(parts, a)                         # persistent-CTA scheduling does not leak into results


@pytest.mark.parametrize("n_lags", [None, 600])          # default window; two lag windows
def test_tc_kernel_missing_data_variant(pkg, ctx, n_lags):
    n_t, n_trials = 5000, 6
    om, mk = _models(pkg, n_t, n_trials, n_lags, missing=True)
    g_tc, g_ff = _gpu_models(mk, ctx)
    rng = np.random.default_rng(4)
    theta = np.array([[0.3], [2.0], [0.07]])
    u0 = rng.standard_normal((3, n_trials)); noise = rng.standard_normal((3, n_trials, n_t))
    d_tc, s_tc = g_tc.distances(theta, u0=u0, noise=noise, return_stats=True)
    d_ff, s_ff = g_ff.distances(theta, u0=u0, noise=noise, return_stats=True)
    assert om.missing_mask.any()
    for i in range(3):
        ac = omodels.summary_stats(om, omodels.generate_data(om, theta[i], u0=u0[i], noise=noise[i]))
        assert np.max(np.abs(s_tc[i] - ac)) <= ACF_TOL
    assert np.max(np.abs(s_tc - s_ff)) <= ACF_TOL


def test_unsupported_shapes_fall_back_to_cuda_core_kernel(pkg, ctx):
    """n_lags beyond four lag windows (1537): the automatic choice is k_abc_acf, tuning.abc_tc = 1 refuses."""
    om, mk = _models(pkg, 2500, 1, 1538)
    assert ctx.get_tuning().abc_tc == -1
    assert mk().gpu_model(ctx).kernel_info()[0] == "k_abc_acf"
    with ctx.tuned(abc_tc=1):
        with pytest.raises(pkg.ItjlError):
            mk().gpu_model(ctx)


def test_single_term_products_and_rounding_correction_over_many_particles(pkg, ctx):
    """C5 shape (50 x 5000 = 2.5e5 samples per particle => hi*hi products only, ONE accumulator segment per
    particle): the summary statistics of 1184 Philox-driven particles against the fp64 C restatement of the
    reference.  The dropped cross terms behave like the documented zero-mean residual, and the
    round-toward-zero of the tensor-core accumulator (-1.1e-5 at lags with ac ~ 1 when uncorrected) is
    removed by the epilogue's first-order correction: no error proportional to the value is left.
    The CUDA-core kernel (fp32 FFMA sums of the same series) is held to the same tolerance."""
    from oracle import cport
    om, mk = _models(pkg, 5000, 50, 414)
    g_tc, g_ff = _gpu_models(mk, ctx)
    plan = (ctypes.c_int32 * 16)()
    assert pkg._lib.lib().itjl_tc_plan(5000, 414, 50, 232448, 148, plan) == 0
    assert plan[6] == 50 and plan[7] == 1                   # seg_trials, n_terms
    rng = np.random.default_rng(7)
    theta = rng.uniform(0.05, 5.0, size=(1184, 1))
    d_tc, s_tc = g_tc.distances(theta, seed=21, generation=4, return_stats=True)
    d_ff, s_ff = g_ff.distances(theta, seed=21, generation=4, return_stats=True)
    d_or, s_or = cport.abc_distances(om, theta, seed=21, generation=4, return_stats=True)
    for name, s, d in (("tc", s_tc, d_tc), ("ffma", s_ff, d_ff)):
        err = s - s_or
        slope = float((s_or * err).sum() / (s_or * s_or).sum())
        print(f"{name} vs fp64 oracle over 1184 particles: max {np.abs(err).max():.3g} rms {np.sqrt(np.mean(err ** 2)):.3g} "
              f"slope {slope:+.2g}")
        assert np.abs(err).max() <= ACF_TOL, (name, np.abs(err).max())
        assert np.sqrt(np.mean(err ** 2)) <= 1.5e-6, name
        assert abs(slope) <= 1e-6, (name, slope)              # uncorrected tensor-core sums: -1.1e-5
        assert np.all(np.abs(d - d_or) <= 2 * np.sqrt(d_or) * ACF_TOL + ACF_TOL ** 2 + 1e-5 * d_or)
    # deterministic: integer fixed-point lag table, fixed segment order
    assert np.array_equal(g_tc.distances(theta[:300], seed=21, generation=4), d_tc[:300])   # also: stats_out does not change a bit


def test_missing_data_model_with_many_trials_against_fp64_oracle(pkg, ctx):
    """Missing-data model at 50 x 5000 samples per particle (FFMA tail tile; the masked samples all carry the
    same value and hence the same fp16 rounding error, so the model keeps the hi*lo term whatever its size):
    32 Philox-driven particles against the fp64 C restatement of comp_ac_time_missing."""
    from oracle import cport
    om, mk = _models(pkg, 5000, 50, 460, missing=True)
    g_tc, _ = _gpu_models(mk, ctx)
    plan = (ctypes.c_int32 * 16)()
    assert pkg._lib.lib().itjl_tc_plan(5000, 460, 50, 232448, 148, plan) == 0 and plan[10] == 1      # one tail tile
    assert g_tc.kernel_info()[1] > 2000                   # two product terms: ~2438 executed tensor flops per sample
    theta = np.random.default_rng(23).uniform(0.05, 5.0, size=(32, 1))
    d_tc, s_tc = g_tc.distances(theta, seed=9, generation=2, return_stats=True)
    d_or, s_or = cport.abc_distances(om, theta, seed=9, generation=2, return_stats=True)
    err = np.abs(s_tc - s_or)
    print(f"tc (missing) vs fp64 oracle over 32 particles: max {err.max():.3g} rms {np.sqrt(np.mean(err ** 2)):.3g}")
    assert err.max() <= ACF_TOL, err.max()
    assert np.sqrt(np.mean(err ** 2)) <= 2e-6
    assert np.all(np.abs(d_tc - d_or) <= 2 * np.sqrt(d_or) * ACF_TOL + ACF_TOL ** 2 + 1e-5 * d_or)


def test_statistics_of_many_philox_particles_against_fp64_oracle(pkg, ctx):
    """C5 shape, 296 Philox-driven particles (two per SM, every simulate group and accumulator segment in
    use): the trial-mean ACF of the tensor-core kernel against the fp64 C restatement of the reference
    (oracle/c, which itself is pinned to the NumPy oracle in test_oracle_c_port.py) -- the device draws
    differ from the oracle's only by the SFU Box-Muller (~5e-7 per draw)."""
    from oracle import cport
    om, mk = _models(pkg, 5000, 50, 414)
    g_tc, _ = _gpu_models(mk, ctx)
    rng = np.random.default_rng(17)
    theta = rng.uniform(0.05, 5.0, size=(296, 1))
    d_tc, s_tc = g_tc.distances(theta, seed=5, generation=3, return_stats=True)
    d_or, s_or = cport.abc_distances(om, theta, seed=5, generation=3, return_stats=True)
    err = np.abs(s_tc - s_or)
    assert err.max() <= ACF_TOL, err.max()
    assert np.sqrt(np.mean(err ** 2)) <= 3e-6
    assert np.all(np.abs(d_tc - d_or) <= 2 * np.sqrt(d_or) * ACF_TOL + ACF_TOL ** 2 + 1e-5 * d_or)
    print(f"tc vs fp64 oracle over 296 particles: max {err.max():.3g} rms {np.sqrt(np.mean(err ** 2)):.3g}")


@pytest.mark.parametrize("n_t,n_trials,n_lags,missing", [(7722, 64, 1208, True), (7397, 64, 1122, False), (5000, 200, 414, False)])
def test_both_kernels_on_many_long_trials_against_fp64_oracle(pkg, ctx, n_t, n_trials, n_lags, missing):
    """Many long trials and few time streams per lag tile (found by benchmarks/tc_fuzz.py): a single fp32
    accumulator per thread for the whole particle made the CUDA-core kernel lose up to 1.3e-4 at the small lags
    (stagnation); it now sums each trial separately.  The tensor-core kernel runs three lag windows / two
    accumulator segments here.  Lag 0 is exactly 1, as in the reference."""
    from oracle import cport
    om, mk = _models(pkg, n_t, n_trials, n_lags, missing=missing, seed=44)
    g_tc, g_ff = _gpu_models(mk, ctx)
    theta = np.random.default_rng(43).uniform(0.05, 5.0, size=(8, 1))
    d_or, s_or = cport.abc_distances(om, theta, seed=3, generation=1, return_stats=True)
    for name, g in (("tc", g_tc), ("ffma", g_ff)):
        d, s = g.distances(theta, seed=3, generation=1, return_stats=True)
        err = np.abs(s - s_or)
        print(f"{name}: n_t={n_t} trials={n_trials} L={n_lags} missing={missing} max {err.max():.3g}")
        assert err.max() <= ACF_TOL, (name, err.max())
        assert np.all(s[:, 0] == 1.0)
        assert np.all(np.abs(d - d_or) <= 2 * np.sqrt(d_or) * ACF_TOL + ACF_TOL ** 2 + 1e-5 * d_or)


def test_tc_kernel_oscillation_acf_variant(pkg, ctx):
    """one_timescale_and_osc_model with the ACF summary (OSC = true template): n_t = 10 000 leaves room
    for two simulate groups only; tensor-core and CUDA-core kernels agree with the oracle."""
    dt, n_t, n_trials = 1 / 500, 10000, 5
    t = dt * np.arange(1, n_t + 1)
    data = oou.generate_ou_with_oscillation([0.05, 10.0, 0.95], dt, n_t * dt, n_trials, 0.0, 1.0, seed=11, n_t=n_t)
    po = [omodels.Normal(.1, .1), omodels.Normal(10, 5), omodels.Uniform(0, 1)]
    om = omodels.one_timescale_and_osc_model(data, t, prior=po, summary_method="acf")
    mk = lambda: pkg.one_timescale_and_osc_model(data, t, "abc", prior=[pkg.Normal(.1, .1), pkg.Normal(10, 5), pkg.Uniform(0, 1)],
                                                 summary_method="acf")
    assert mk().n_lags == om.n_lags <= 509
    g_tc, g_ff = _gpu_models(mk, ctx)
    rng = np.random.default_rng(2)
    theta = np.array([[0.05, 10.0, 0.95], [0.1, 12.0, 0.5], [0.02, 8.0, 0.3]])
    u0 = rng.standard_normal((3, n_trials)); noise = rng.standard_normal((3, n_trials, n_t))
    phases = rng.uniform(0, 2 * np.pi, size=(3, n_trials))
    d_tc, s_tc = g_tc.distances(theta, u0=u0, noise=noise, phases=phases, return_stats=True)
    d_ff, s_ff = g_ff.distances(theta, u0=u0, noise=noise, phases=phases, return_stats=True)
    for i in range(3):
        ac = omodels.summary_stats(om, omodels.generate_data(om, theta[i], u0=u0[i], noise=noise[i], phases=phases[i]))
        assert np.max(np.abs(s_tc[i] - ac)) <= 1e-5
    assert np.max(np.abs(s_tc - s_ff)) <= 1e-5


def e_shared(n_trials):
    """What both fused kernels share against the fp64 reference: fp32 simulation, centring, conversion and fp32
    accumulation of the lag sums: <= 1.3e-6 measured from tau = T / 100 to 50 T, 1 to 200 trials (it was up to 1.2e-5
    for single near-unit-root trials until the per-chunk sums of the carry powers moved to fp64, simulate.cuh)."""
    return 1.5e-6


@pytest.mark.parametrize("n_trials", [1, 2, 7, 50, 200])
def test_tensor_core_kernel_stays_within_its_predicted_error(pkg, ctx, n_trials):
    """The automatic plan (tuning.abc_tc = -1: product terms and segments from the error model,
    itjl_tc_predicted_error) over timescales from 0.01 to 50 times the trial length, at 1 .. 200 trials of 5000
    samples.  The model bounds what the tensor-core lag sums ADD to the error of the fp32 path (default bound
    5e-6); the CUDA-core kernel shows the shared part (E_SHARED); together they stay within the contract's 1e-5."""
    from oracle import cport
    n_t, n_lags = 5000, 414
    om, mk = _models(pkg, n_t, n_trials, n_lags)
    ctx.reset_tuning()
    g_auto = mk().gpu_model(ctx)
    with ctx.tuned(abc_tc=0):
        g_ff = mk().gpu_model(ctx)
    plan = (ctypes.c_int32 * 16)()
    assert pkg._lib.lib().itjl_tc_plan(n_t, n_lags, n_trials, 232448, 148, plan) == 0 and plan[15] == 0
    assert g_auto.kernel_info()[0] == "k_abc_acf_tc"
    pred = pkg._lib.tc_predicted_error(n_t, n_trials, plan[7], plan[6] * 40)
    assert pred <= 5e-6
    T = n_t * DT
    taus = T * np.array([0.01, 0.02, 0.05, 0.1, 0.2, 0.5, 1.0, 2.0, 5.0, 10.0, 20.0, 50.0])
    theta = np.repeat(taus, 8 if n_trials <= 7 else 2).reshape(-1, 1)
    d_or, s_or = cport.abc_distances(om, theta, seed=31, generation=5, return_stats=True)
    d_tc, s_tc = g_auto.distances(theta, seed=31, generation=5, return_stats=True)
    d_ff, s_ff = g_ff.distances(theta, seed=31, generation=5, return_stats=True)
    e_tc, e_ff = np.abs(s_tc - s_or).max(axis=1), np.abs(s_ff - s_or).max(axis=1)
    by_tau = lambda e: " ".join(f"{x:.1e}" for x in e.reshape(len(taus), -1).max(axis=1))
    print(f"n_trials={n_trials} terms={plan[7]} seg={plan[6]} predicted {pred:.2e}\n  tc   by tau/T: {by_tau(e_tc)}\n"
          f"  ffma by tau/T: {by_tau(e_ff)}\n  |tc - ffma| max {np.abs(s_tc - s_ff).max():.2e}")
    E_SHARED = e_shared(n_trials)
    assert e_ff.max() <= E_SHARED, e_ff.max()
    assert np.abs(s_tc - s_ff).max() <= pred + 1.5e-6          # + the FFMA kernel's own fp32 summation
    assert e_tc.max() <= pred + E_SHARED <= 1e-5
    tol = pred + E_SHARED
    assert np.all(np.abs(d_tc - d_or) <= 2 * np.sqrt(d_or) * tol + tol ** 2 + 1e-5 * d_or)
