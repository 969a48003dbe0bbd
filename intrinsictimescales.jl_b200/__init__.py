"""itjl_b200 - B200 (sm_100a) implementation of the data-parallel hot path of
IntrinsicTimescales.jl: the adaptive-ABC inner loop (simulate OU trials -> ACF / NaN-aware
ACF / PSD summary -> distance -> accept) and the batched acw() reducers.

This package is the host-side mirror of the reference's public interface for that path
(same names and argument meaning as the Julia functions; Julia is not available in this
image - the Julia `ccall` shim is documented in INTEGRATION.md).  All numerics run in
hand-written CUDA kernels behind the C ABI of include/itjl_b200.h (libitjl_b200.so);
there is no CPU fallback.
"""
from . import _lib
from ._lib import Context, ItjlError, default_context, library_path
from .abc import (ABCResults, abc_results, basic_abc, calc_weights, compute_adaptive_alpha,
                  draw_theta_pmc, draw_theta_pmc_device, effective_sample_size, find_MAP, get_param_dict_abc, int_fit,
                  pmc_abc, posterior_predictive, select_epsilon, weighted_covar)
from .acw import ACWResults, DeviceData, acw, possible_acwtypes, upload
from .models import (Normal, Uniform, TimescaleModel, check_model_inputs, distance_function,
                     draw_theta, generate_data, generate_data_and_reduce, one_timescale_and_osc_model,
                     one_timescale_and_osc_with_missing_model,
                     one_timescale_model, one_timescale_with_missing_model, summary_stats)
from .ou import generate_ou_process, generate_ou_with_oscillation, generate_two_timescale
from .summary import (comp_ac_fft, comp_ac_time_missing, comp_psd, comp_psd_adfriendly, comp_psd_lombscargle,
                      lombscargle_autofrequency, nextfastfft)
from .utils import (acw0, acw50, acw50_analytical, acweuler, expdecay, find_knee_frequency, find_oscillation_peak,
                    fit_expdecay, fit_expdecay_3_parameters, fooof_fit, knee_from_tau, lorentzian,
                    lorentzian_initial_guess, tau_from_acw50, tau_from_knee)

__version__ = "0.1.0"
