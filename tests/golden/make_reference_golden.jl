# Generates tests/golden/reference_v1.json: outputs of the UNMODIFIED reference
# (duodenum96/IntrinsicTimescales.jl v0.7.0) on the committed inputs of
# tests/golden/reference_inputs_v1/ (written by make_reference_inputs.py).
#
#   julia --project=/path/to/IntrinsicTimescales.jl tests/golden/make_reference_golden.jl
#
# Only Base + the reference package are used (inputs are raw Float64 files, the JSON is emitted by
# hand), so no extra Julia package is needed.  Every key emitted here is compared by
# tests/test_reference_golden.py with the CPU oracle (CPU test) and with the CUDA path (GPU test).
# Julia is not installed in the build image: this script has not been executed there; the loader
# tests skip until reference_v1.json exists.
#
# Reference functions exercised (file:line in the reference checkout):
#   comp_ac_fft              src/stats/summary.jl:46-63, 79-89
#   comp_ac_time_missing     src/stats/summary.jl:455-496
#   comp_psd_adfriendly      src/stats/summary.jl:183-235
#   comp_psd (periodogram)   src/stats/summary.jl:111-160       (DSP.jl periodogram, Hamming)
#   acw                      src/core/acw.jl:109-307
#   fit_expdecay             src/utils/utils.jl:113-138         (NonlinearSolve LevenbergMarquardt)
#   acw_romberg              src/utils/utils.jl:332-345         (Romberg.jl)
#   acw0 / acw50 / acweuler  src/utils/utils.jl:218-302
#   select_epsilon           src/core/abc.jl:700-761            (StatsBase.percentile)
#   calc_weights (2 params)  src/core/abc.jl:547-566
#   weighted_covar           src/core/abc.jl:582-613
#   effective_sample_size    src/core/abc.jl:632-642
#   linear_distance / logarithmic_distance  src/stats/distances.jl:29-53
#   comp_psd_lombscargle     src/stats/summary.jl:315-360       (LombScargle.jl, autofrequency grid)
#   lorentzian_initial_guess / find_knee_frequency / find_oscillation_peak / fooof_fit
#                            src/utils/utils.jl:423-445, 479-538, 721-765, 598-675
#   tau_from_knee / knee_from_tau            src/utils/utils.jl:197-198
#   fit_expdecay_3_parameters                src/utils/utils.jl:159-192
#   acw(:knee) / acw(:tau; skip_zero_lag)    src/core/acw.jl:212-216, 242-270

using IntrinsicTimescales
import IntrinsicTimescales.ABC as ABCmod
import Distributions
import StatsBase
import Statistics

const HERE = @__DIR__
const INDIR = joinpath(HERE, "reference_inputs_v1")
const FS = 500.0
const DT = 1.0 / FS

function load_inputs()
    inputs = Dict{String, Array{Float64}}()
    for line in eachline(joinpath(INDIR, "manifest.txt"))
        parts = split(line)
        isempty(parts) && continue
        name = String(parts[1])
        dims = Tuple(parse.(Int, parts[2:end]))
        a = Array{Float64}(undef, dims...)
        open(joinpath(INDIR, name * ".f64"), "r") do io
            read!(io, a)
        end
        inputs[name] = a
    end
    return inputs
end

# ---- minimal JSON emitter: {"key": {"shape": [...], "data": [...column-major...]}, ...} ----
fmt(x::Real) = isnan(x) ? "NaN" : (isinf(x) ? (x > 0 ? "Infinity" : "-Infinity") : repr(Float64(x)))
fmt(x::Integer) = string(x)

const ENTRIES = String[]

function emit(key::String, value)
    if value isa Number
        push!(ENTRIES, "\"$key\": {\"shape\": [], \"data\": [$(fmt(value))]}")
    else
        a = collect(value)
        shape = join(string.(size(a)), ", ")
        data = join([fmt(v) for v in vec(a)], ", ")
        push!(ENTRIES, "\"$key\": {\"shape\": [$shape], \"data\": [$data]}")
    end
    return nothing
end

nan_if_nothing(x) = x === nothing ? NaN : x
as_array(v) = v isa Number ? [Float64(v)] : Float64.(collect(v))

function main()
    inp = load_inputs()
    x = inp["x_complete"]
    xm = inp["x_missing"]
    xo = inp["x_osc"]
    xw = inp["x_white"]
    acf_rows = inp["acf_rows"]

    # ---- summary statistics (dims = last = time) ----
    emit("comp_ac_fft", comp_ac_fft(x; n_lags=60))
    emit("comp_ac_fft_full_row1", comp_ac_fft(x[1, :]))
    emit("comp_ac_time_missing", comp_ac_time_missing(xm; n_lags=60))
    psd, freqs = comp_psd_adfriendly(xo, FS)
    emit("comp_psd_adfriendly_power", psd)
    emit("comp_psd_adfriendly_freqs", freqs)
    psd2, freqs2 = comp_psd(xo, FS)
    emit("comp_psd_periodogram_power", psd2)
    emit("comp_psd_periodogram_freqs", freqs2)

    # ---- acw: every ACF-based type, complete / missing / white-noise / averaged ----
    types = [:acw0, :acw50, :acweuler, :auc, :tau]
    for (tag, data, kwargs) in (("ou", x, (;)), ("white", xw, (;)),
                                ("missing", xm, (;)), ("ou_avg", x, (; average_over_trials=true)),
                                ("ou_nlags40", x, (; n_lags=40)))
        r = acw(copy(data), FS; acwtypes=types, kwargs...)
        for (i, t) in enumerate(types)
            emit("acw_$(tag)_$(t)", Float64.(nan_if_nothing.(collect(r.acw_results[i]))))
        end
        emit("acw_$(tag)_n_lags", r.n_lags)
        emit("acw_$(tag)_acf", r.acf)
        emit("acw_$(tag)_lags", collect(r.lags))
    end

    # ---- reducers on given ACF rows ----
    lags = collect(DT .* (0:(size(acf_rows, 2) - 1)))
    emit("fit_expdecay", Float64.(fit_expdecay(lags, acf_rows)))
    emit("acw0_rows", Float64.(acw0(lags, acf_rows)))
    emit("acw50_rows", Float64.(acw50(lags, acf_rows)))
    emit("acweuler_rows", Float64.(acweuler(lags, acf_rows)))
    for k in 2:17                                   # Romberg.jl on 2..17 points (every factorisation of k-1 <= 16)
        emit("acw_romberg_$(k)", [acw_romberg(DT, acf_rows[i, 1:k]) for i in 1:size(acf_rows, 1)])
    end

    # ---- ABC bookkeeping ----
    d = inp["distances"][:]
    emit("percentile_25_50_75", [StatsBase.percentile(d[.!isnan.(d) .& (d .< 10.0)], p) for p in (25.0, 50.0, 75.0)])
    emit("select_epsilon_iter1", ABCmod.select_epsilon(d, 1.0))
    emit("select_epsilon_high_acc", ABCmod.select_epsilon(d, 0.2; current_acc_rate=0.3, iteration=3, total_iterations=30))
    emit("select_epsilon_low_acc", ABCmod.select_epsilon(d, 0.05; current_acc_rate=0.001, iteration=7, total_iterations=30))
    emit("select_epsilon_in_band", ABCmod.select_epsilon(d, 0.05; current_acc_rate=0.0105, iteration=7, total_iterations=30))
    emit("compute_adaptive_alpha", [ABCmod.compute_adaptive_alpha(it, acc, 0.01; total_iterations=30)
                                    for (it, acc) in ((1, 0.5), (10, 0.011), (20, 0.02), (30, 0.001))])
    prior = [Distributions.Uniform(0.0, 2.0), Distributions.Normal(10.0, 5.0)]
    w_new = ABCmod.calc_weights(inp["theta_prev"], inp["theta_new"], inp["tau_squared"], inp["weights_prev"][:], prior)
    emit("calc_weights_2param", w_new)
    emit("weighted_covar", weighted_covar(inp["theta_new"], w_new))
    emit("effective_sample_size", effective_sample_size(w_new))

    # ---- PSD side: Lomb-Scargle, knee / peak / FOOOF fits, knee-based acw, 3-parameter decay ----
    times = collect(DT:DT:(size(xm, 2) * DT))
    ls_power, ls_freqs = comp_psd_lombscargle(times, xm, isnan.(xm), DT)
    emit("comp_psd_lombscargle_power", ls_power)
    emit("comp_psd_lombscargle_freqs", collect(ls_freqs))
    pc, fc = comp_psd(x, FS)
    mean_c = vec(Statistics.mean(pc, dims=1))
    mean_o = vec(Statistics.mean(psd2, dims=1))
    fo = collect(freqs2)
    fc = collect(fc)
    for (ptag, m, f) in (("ou", mean_c, fc), ("osc", mean_o, fo))
        emit("lorentzian_initial_guess_$(ptag)", lorentzian_initial_guess(m, f; min_freq=f[1], max_freq=f[end]))
        emit("find_knee_frequency_$(ptag)", find_knee_frequency(m, f; min_freq=f[1], max_freq=f[end]))
        emit("find_knee_frequency_band_$(ptag)", find_knee_frequency(m, f; min_freq=1.0, max_freq=100.0))
    end
    emit("find_oscillation_peak_osc", find_oscillation_peak(mean_o, fo; min_freq=2.0, max_freq=40.0))
    emit("find_oscillation_peak_none", find_oscillation_peak(collect(1.0 ./ (1.0 .+ (fc ./ 3.0) .^ 2)), fc; min_freq=2.0, max_freq=40.0))
    emit("fooof_fit_osc_knee", fooof_fit(mean_o, fo; min_freq=1.0, max_freq=100.0, oscillation_peak=true, max_peaks=2,
                                         return_only_knee=true))
    emit("fooof_fit_ou_knee_only", fooof_fit(mean_c, fc; min_freq=1.0, max_freq=100.0, oscillation_peak=false,
                                             return_only_knee=true))
    emit("tau_from_knee", tau_from_knee([0.5, 3.0, 40.0]))
    emit("knee_from_tau", knee_from_tau([0.004, 0.05, 0.3]))
    emit("acw_knee_ou_avg", as_array(acw(copy(x), FS; acwtypes=:knee, average_over_trials=true, oscillation_peak=false).acw_results))
    emit("acw_knee_osc_avg", acw(copy(xo), FS; acwtypes=:knee, average_over_trials=true, freqlims=(1.0, 100.0),
                                 oscillation_peak=true, max_peaks=2).acw_results |> as_array)
    emit("acw_knee_missing_avg", acw(copy(xm), FS; acwtypes=:knee, average_over_trials=true, freqlims=(1.0, 100.0),
                                     oscillation_peak=false).acw_results |> as_array)
    emit("fit_expdecay_3_parameters", Float64.(fit_expdecay_3_parameters(lags, acf_rows)))
    emit("acw_tau_skip_zero_lag", as_array(acw(copy(x), FS; acwtypes=:tau, skip_zero_lag=true).acw_results))

    # ---- distances ----
    emit("linear_distance", linear_distance(inp["dist_a"][:], inp["dist_b"][:]))
    emit("logarithmic_distance", logarithmic_distance(inp["dist_a"][:], inp["dist_b"][:]))

    open(joinpath(HERE, "reference_v1.json"), "w") do io
        write(io, "{\n" * join(ENTRIES, ",\n") * "\n}\n")
    end
    println("wrote ", length(ENTRIES), " entries to ", joinpath(HERE, "reference_v1.json"))
end

main()
