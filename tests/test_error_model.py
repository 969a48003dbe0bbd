"""The error model behind the automatic kernel choice (itjl_tc_predicted_error / itjl_tc_plan, host only: no GPU).
tests/test_gpu_tc.py::test_tensor_core_kernel_stays_within_its_predicted_error holds the kernel to it."""
import ctypes

import numpy as np


def plan(pkg, n_t, n_lags, n_trials):
    out = (ctypes.c_int32 * 16)()
    rc = pkg._lib.lib().itjl_tc_plan(n_t, n_lags, n_trials, 232448, 148, out)
    return rc, dict(zip("groups chunk ksteps rows_total used_cols mixed seg_trials n_terms tc_lags tail_start tail_lt n_ops "
                        "smem_bytes grid n_windows to_cuda_core".split(), list(out)))


def test_predicted_error_is_monotone_and_scales_as_documented(pkg):
    e = pkg._lib.tc_predicted_error
    # more product terms, more samples, shorter segments: never worse
    for n_t, n_trials in [(5000, 2), (5000, 50), (600, 3), (10000, 100)]:
        assert e(n_t, n_trials, 1, 40) >= e(n_t, n_trials, 2, 40) >= e(n_t, n_trials, 3, 40)
        assert e(n_t, n_trials, 1, 2000) > e(n_t, n_trials, 1, 40)
        assert e(n_t, 4 * n_trials, 1, 40) < e(n_t, n_trials, 1, 40)
    # dropped cross terms: rms ~ 1 / sqrt(samples)
    a = e(5000, 50, 1, 0) - e(5000, 50, 3, 0)
    b = e(5000, 200, 1, 0) - e(5000, 200, 3, 0)
    assert a > b > 0
    # masked samples share one rounding residual: the oscillation + missing-data model (masked samples parked at
    # -data_mean) cannot drop cross terms, the plain missing-data model can keep hi*lo only
    assert e(5000, 100, 1, 2000, 0.1, 2.0) > 5e-5 and e(5000, 100, 3, 2000, 0.1, 2.0) < 5e-6
    assert e(5000, 100, 2, 2000, 0.1, 0.0) < 5e-6


def test_plans_chosen_from_the_error_model(pkg):
    rc, p = plan(pkg, 5000, 414, 50)                       # BASELINE C5: hi*hi alone, one segment
    assert rc == 0 and p["n_terms"] == 1 and p["seg_trials"] == 50 and p["to_cuda_core"] == 0
    rc, p = plan(pkg, 5000, 297, 2)                        # C1: 1e4 samples per particle -> all three terms
    assert rc == 0 and p["n_terms"] == 3
    rc, p = plan(pkg, 5000, 414, 20)                       # 1e5 samples
    assert rc == 0 and p["n_terms"] in (1, 2)
    t = [plan(pkg, 5000, 414, n)[1]["n_terms"] for n in (1, 2, 5, 10, 20, 50, 100, 400)]
    assert t == sorted(t, reverse=True)                    # never more terms for more samples
    rc, p = plan(pkg, 5000, 1538, 3)
    assert rc != 0                                         # beyond four lag windows: no plan at all


def test_tuning_struct_round_trip_without_a_gpu(pkg):
    t = pkg._lib.Tuning()
    assert pkg._lib.lib().itjl_tuning_default(ctypes.byref(t)) == 0
    assert (t.abc_tc, t.tc_terms, t.tc_tail_max, t.tc_bias, t.acw_stream, t.pinned_upload, t.pmc_fp32) == (-1, 0, -1, 1, 1, 1, -1)
    assert t.tc_error_bound == 5e-6
    assert ctypes.sizeof(pkg._lib.Tuning) == 64            # 14 x int32 + double: the layout INTEGRATION.md shows for Julia
