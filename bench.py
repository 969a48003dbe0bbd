#!/usr/bin/env python
"""Benchmark of the aABC hot path: OU samples simulated + ACF'd + scored per second.

Headline workload (BASELINE.json config C5): one "step" = one ABC generation of 10^6 particles, each
simulating 50 trials x 5000 samples (dt = 0.002) of an OU process, reducing them to the trial-mean ACF
(n_lags set by the synthetic data exactly as one_timescale_model does) and scoring the linear distance;
particles are sharded over the ranks (strong scaling: the 10^6 particles are fixed, each of N ranks takes a
contiguous 1/N) and the per-particle distances are all-gathered INSIDE the library (NCCL, device to device).

  python bench.py --gpus 1 --steps K --warmup W                   (our arm, CUDA kernels)
  python bench.py --impl reference --gpus 1 --steps K --warmup W  (CPU restatement of the reference's algorithm)
  python bench.py --config c5sweep ...                            (the value is the 3-generation sweep instead)
One JSON line on stdout (rank 0).  At N = 1 the line also carries `sub`: BASELINE configs C2 (acw), C3 (missing
data), C4 (oscillation + PSD) and the 3-generation C5 sweep, each with its own roofline / cpu_baseline / e2e.
"""
import argparse
import json
import os
import re
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

DT = 0.002
N_T = 5000
N_TRIALS = 50
TAU_TRUE = 0.3
PRIOR = (0.05, 5.0)
METRIC = "aabc_ou_samples_simulated_acf_scored_per_sec"
UNIT = "samples/s"


# ------------------------------------------------------------- synthetic observed data (NumPy only) ----
def ou_rows(rng, n_trials, n_t, tau, dt):
    """Exact OU transition, standardised per trial (no GPU, no oracle)."""
    a = np.exp(-dt / tau)
    s = np.sqrt(tau / 2 * (1 - a * a))
    x = np.empty((n_trials, n_t))
    prev = rng.standard_normal(n_trials)
    z = rng.standard_normal((n_t, n_trials))
    for k in range(n_t):
        prev = a * prev + s * z[k]
        x[:, k] = prev
    return (x - x.mean(axis=1, keepdims=True)) / x.std(axis=1, ddof=1, keepdims=True)


def observed_data_host():
    """C5: OU tau = 0.3, 50 x 5000, seed 5."""
    return ou_rows(np.random.default_rng(5), N_TRIALS, N_T, TAU_TRUE, DT), DT * np.arange(1, N_T + 1)


def observed_c3():
    """C3: OU tau = 0.3, 100 x 5000, one contiguous 10 % NaN gap per trial, seed 7."""
    rng = np.random.default_rng(7)
    x = ou_rows(rng, 100, 5000, 0.3, DT)
    for r in range(100):
        s = int(rng.integers(0, 4500))
        x[r, s:s + 500] = np.nan
    return x, DT * np.arange(1, 5001)


def observed_c4():
    """C4: generate_ou_with_oscillation([0.05, 10, 0.95], 1/500, 20 s, 100 trials, 0, 1), seed 11."""
    rng = np.random.default_rng(11)
    dt, n_t = 1 / 500, 10000
    t = dt * np.arange(1, n_t + 1)
    ou = ou_rows(rng, 100, n_t, 0.05, dt) * np.sqrt(0.05 / 2)            # raw OU: variance tau / 2
    osc = np.sqrt(2.0) * np.sin(rng.uniform(0, 2 * np.pi, size=(100, 1)) + 2 * np.pi * 10.0 * t[None, :])
    y = np.sqrt(0.05) * osc + np.sqrt(0.95) * ou
    return (y - y.mean(axis=1, keepdims=True)) / y.std(axis=1, ddof=1, keepdims=True), t


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled while the timed region runs."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for nm, v in zip(names, f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": float(max(mx)) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------ CPU arm ----
def cpu_rate(om, n_particles, threads, seed=1, prior=PRIOR, n_theta=1, theta=None):
    """OU samples/s of the C restatement (oracle/c) on `threads` host threads."""
    from oracle import cport
    if theta is None:
        theta = np.random.default_rng(seed).uniform(*prior, size=(n_particles, n_theta))
    t0 = time.perf_counter()
    cport.abc_distances(om, theta, seed=seed, generation=1, n_threads=threads)
    dt = time.perf_counter() - t0
    return theta.shape[0] * om.numTrials * om.n_t / dt, dt


def oracle_model():
    from oracle import models as omodels
    data, t = observed_data_host()
    return omodels.one_timescale_model(data, t, prior=[omodels.Uniform(*PRIOR)])


CPU_NOTE = ("C restatement of the reference algorithm (exact OU transition + FFT ACF in fp64, vectorised radix-2 FFT with "
            "two real trials per complex transform); Julia is not installed on this image")


def cpu_baseline(target_seconds=12.0):
    from oracle import cport
    om = oracle_model()
    cores = cport.max_threads()
    rate, _ = cpu_rate(om, cores, cores, seed=0)                      # calibration
    n = max(cores, int(target_seconds * rate / (N_TRIALS * N_T)))
    rate, secs = cpu_rate(om, n, cores)
    rate1, secs1 = cpu_rate(om, 12, 1, seed=3)
    return {"value": rate, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{n} particles x {N_TRIALS} trials x {N_T} samples (n_lags={om.n_lags}), {CPU_NOTE}, "
                      f"{cores} pthreads, {secs:.1f} s",
            "one_thread": {"value": rate1, "unit": UNIT, "cores": 1,
                           "sample": f"12 particles on one thread, {secs1:.1f} s: what the reference's serial loop gets"}}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import cport
    om = oracle_model()
    cores = cport.max_threads()
    rate, _ = cpu_rate(om, cores, cores, seed=0)
    per_step = max(cores, int(4.0 * rate / (N_TRIALS * N_T)))          # ~4 s of CPU work per step
    for w in range(args.warmup):
        cpu_rate(om, per_step, cores, seed=100 + w)
    t0 = time.perf_counter()
    for k in range(args.steps):
        cpu_rate(om, per_step, cores, seed=200 + k)
    secs = time.perf_counter() - t0
    value = args.steps * per_step * N_TRIALS * N_T / secs
    rate1, _ = cpu_rate(om, 12, 1, seed=3)
    sample = (f"{per_step} particles/step x {N_TRIALS} trials x {N_T} samples, n_lags={om.n_lags}; "
              f"{CPU_NOTE}; the reference's serial loop body run on {cores} host threads")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * secs / args.steps,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": "C5 aABC generation: 50 trials x 5000 samples per particle, ACF n_lags="
                                   f"{om.n_lags}, linear distance", "particles_per_step": per_step},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
                             "one_thread": {"value": rate1, "unit": UNIT, "cores": 1}},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------ GPU arm ----
def measured_peak(key, fallback):
    """Roofline denominator: MEASURED_PEAKS.json (driver-written) or the profiling guide's fallback."""
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            v = float(json.load(f)[key])
        return v, f"MEASURED_PEAKS.json {key}"
    except Exception:
        return float(fallback), f"fallback ({key} absent): B200_PROFILING.md figure"


def ncu_traffic(summary_file):
    """dram__bytes_read.sum + dram__bytes_write.sum of the committed `ncu --set full` summary (profiles/), bytes."""
    unit = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    try:
        txt = open(os.path.join(ROOT, "profiles", summary_file)).read()
        tot = 0.0
        for key in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            m = re.search(re.escape(key) + r" \[(\w+)\] = ([0-9.eE+-]+)", txt)
            tot += float(m.group(2)) * unit[m.group(1)]
        return tot
    except Exception:
        return None


def event_ms(torch, stream, fn, reps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    stream.synchronize()
    e0.record(stream)
    for _ in range(reps):
        fn()
    e1.record(stream)
    stream.synchronize()
    return e0.elapsed_time(e1) / reps


def sub_c2(pkg, ctx, torch, stream, hbm_peak, hbm_src, cpu):
    """BASELINE C2: acw() with all ACF-based acwtypes on randn(10000 x 5000) Float64, fs = 100."""
    types = ["acw0", "acw50", "acweuler", "auc", "tau"]
    d = np.asfortranarray(np.random.default_rng(7).standard_normal((10000, 5000)))
    pkg.acw(d, 100.0, acwtypes=types, ctx=ctx)                         # warm-up (allocations, pinned ring)
    t0 = time.perf_counter()
    reps = 3
    for _ in range(reps):
        r = pkg.acw(d, 100.0, acwtypes=types, ctx=ctx)
    e2e_s = (time.perf_counter() - t0) / reps
    dd = pkg.upload(d, ctx=ctx)
    pkg.acw(dd, 100.0, acwtypes=types)
    ctx.timing_reset()
    t0 = time.perf_counter()
    for _ in range(reps):
        pkg.acw(dd, 100.0, acwtypes=types)
    res_s = (time.perf_counter() - t0) / reps
    n_k, k_ms = ctx.timing_get()
    ctx.timing_stop()
    pass_ms = k_ms / max(n_k, 1)                                       # k_acw_stream + k_acw_finish, CUDA events
    bytes_alg = 8.0 * d.size
    out_bytes = 10000 * (3 * 4 + 2 * 8) + 8 * 10000 * r.n_lags
    rec = {"metric": "acw_slices_per_sec", "unit": "slices/s", "value": 10000 / res_s,
           "config": {"workload": "C2 acw(randn(10000 x 5000), fs=100; acw0, acw50, acweuler, auc, tau)", "n_lags": int(r.n_lags),
                      "value_is": "data resident on the device (upload once, repeated acw calls)"},
           "ms_per_call_resident": 1e3 * res_s,
           "roofline": {"bound": "hbm", "kernel": "k_acw_stream + k_acw_finish", "achieved": bytes_alg / (pass_ms * 1e-3) / 1e9,
                        "peak": hbm_peak, "unit": "GB/s", "frac": bytes_alg / (pass_ms * 1e-3) / 1e9 / hbm_peak,
                        "avg_pass_ms": pass_ms, "algorithmic_bytes": bytes_alg, "peak_source": hbm_src,
                        "traffic": ncu_traffic("r02_k_acw_stream_ncu_summary.txt")},
           "e2e": {"value": 10000 / e2e_s, "unit": "slices/s", "ms_per_call": 1e3 * e2e_s,
                   "h2d_bytes_per_step": int(bytes_alg), "d2h_bytes_per_step": int(out_bytes)}}
    if cpu:
        from oracle import acw as oacw
        n = 400
        t0 = time.perf_counter()
        oacw.acw(d[:n], 100.0, acwtypes=types)
        s = time.perf_counter() - t0
        rec["cpu_baseline"] = {"value": n / s, "unit": "slices/s", "cores": 1, "kind": "port",
                               "sample": f"{n} slices, NumPy restatement of acw() (pocketfft ACF + reducers), one thread, {s:.1f} s"}
    return rec


def sub_knee(pkg, ctx, cpu):
    """f3: acw(data, fs; acwtypes = :knee) - periodogram of every slice + FOOOF-style knee fit, 2000 x 5000."""
    fs, n = 100.0, 2000
    d = np.asfortranarray(ou_rows(np.random.default_rng(8), n, 5000, 0.3, 1.0 / fs))
    pkg.acw(d, fs, acwtypes="knee", ctx=ctx)                 # warm-up at full size: the context's buffers grow here
    n0 = ctx.launch_count
    t0 = time.perf_counter()
    r = pkg.acw(d, fs, acwtypes="knee", ctx=ctx)
    s = time.perf_counter() - t0
    tau = np.asarray(r.acw_results)
    rec = {"metric": "acw_knee_slices_per_sec", "unit": "slices/s", "value": n / s,
           "config": {"workload": "acw(OU tau = 0.3 s, 2000 x 5000, fs = 100; acwtypes = :knee, oscillation_peak = true, max_peaks = 1)",
                      "n_freq": int(r.freqs.size), "median_tau": float(np.median(tau)), "kernels": "k_pgram_data + k_transpose + k_knee_fit",
                      "launches": int(ctx.launch_count - n0)},
           "e2e": {"value": n / s, "unit": "slices/s", "ms_per_call": 1e3 * s, "h2d_bytes_per_step": int(d.nbytes),
                   "d2h_bytes_per_step": int(8 * n + 8 * n * r.freqs.size)},
           "roofline": None}
    if cpu:
        from oracle import knee as oknee, summary as osum
        t0 = time.perf_counter()
        psd, fr = osum.comp_psd_periodogram(d[:3], fs)
        for p_ in psd:
            oknee.fooof_fit(p_, fr, fr[0], fr[-1], True, 1, True)
        s1 = time.perf_counter() - t0
        rec["cpu_baseline"] = {"value": 3 / s1, "unit": "slices/s", "cores": 1, "kind": "port",
                               "sample": f"3 slices, NumPy / SciPy oracle (pocketfft periodogram + Nelder-Mead L1 fits), {s1:.1f} s"}
    return rec


def time_model(pkg, ctx, torch, stream, gm, theta, seed):
    """(resident ms per launch from CUDA events on the context stream, e2e seconds per call with host buffers)."""
    P, p = theta.shape
    dev = torch.device("cuda", ctx.device)
    th_dev = torch.from_numpy(np.asfortranarray(theta).reshape(-1, order="F").copy()).to(dev)
    out_dev = torch.empty(P, dtype=torch.float64, device=dev)
    gen = [0]

    def launch():
        gen[0] += 1
        gm.distances_dev(th_dev.data_ptr(), P, p, out_dev.data_ptr(), seed=seed, generation=gen[0])
    for _ in range(2):
        launch()
    ms = event_ms(torch, stream, launch, 3)
    gm.distances(theta, seed=seed, generation=1)
    t0 = time.perf_counter()
    for k in range(2):
        gm.distances(theta, seed=seed, generation=2 + k)
    return ms, (time.perf_counter() - t0) / 2


def sub_model(pkg, ctx, torch, stream, name, model, om_fn, theta, fp32_peak, tensor_peak, tensor_src, cpu, seed=11):
    gm = model.gpu_model(ctx)
    flops, spp = gm.work()
    kname, tflops_exec = gm.kernel_info()
    P = theta.shape[0]
    ms, e2e_s = time_model(pkg, ctx, torch, stream, gm, theta, seed)
    samples = P * spp
    achieved = flops * samples / (ms * 1e-3) / 1e12
    if kname == "k_abc_acf_tc":
        roof = {"bound": "tensor", "kernel": kname, "achieved": achieved, "peak": tensor_peak, "unit": "TFLOP/s",
                "frac": achieved / tensor_peak, "peak_source": tensor_src, "flops_per_sample": flops,
                "executed_tensor_flops_per_sample": tflops_exec, "traffic": None,
                "traffic_note": "compute bound: theta in, one distance out per particle; the trajectories stay in shared memory"}
    else:
        roof = {"bound": "fp32", "kernel": kname, "achieved": achieved, "peak": fp32_peak, "unit": "TFLOP/s",
                "frac": achieved / fp32_peak, "peak_source": "itjl_fp32_peak microbenchmark of this run",
                "flops_per_sample": flops,
                # per launch of the profiled run (profiles/README.md names its particle count), parsed from the committed summary
                "traffic": ncu_traffic({"k_abc_psd2": "r02_k_abc_psd2_ncu_summary.txt", "k_abc_psd": "r01_k_abc_psd_v5_ncu_summary.txt",
                                        "k_abc_pgram": "r02_k_abc_pgram_ncu_summary.txt"}.get(kname, "none")),
                "traffic_note": "compute bound (simulate + FFT in shared memory); DRAM traffic = theta, the bin table and the window"}
    rec = {"metric": METRIC, "unit": UNIT, "value": samples / (ms * 1e-3), "ms_per_launch": ms,
           "config": {"workload": name, "particles_per_launch": P, "n_trials": int(model.numTrials), "n_t": int(model.n_t),
                      "n_stats": int(gm.n_stats)},
           "roofline": roof,
           "e2e": {"value": samples / e2e_s, "unit": UNIT, "h2d_bytes_per_step": int(theta.nbytes), "d2h_bytes_per_step": int(8 * P)}}
    if cpu and kname == "k_abc_pgram":
        # the C port has no mixed-radix periodogram: NumPy oracle (pocketfft), one thread, two particles
        from oracle import models as omodels
        om = om_fn()
        t0 = time.perf_counter()
        for i in range(2):
            omodels.generate_data_and_reduce(om, theta[i], seed=seed, generation=1, particle=i)
        s1 = time.perf_counter() - t0
        rec["cpu_baseline"] = {"value": 2 * spp / s1, "unit": UNIT, "cores": 1, "kind": "port",
                               "sample": f"2 particles of the same model, NumPy oracle (exact OU transition + pocketfft periodogram), {s1:.1f} s"}
    elif cpu:
        from oracle import cport
        om = om_fn()
        cores = cport.max_threads()
        n = 4 * cores
        rate, secs = cpu_rate(om, n, cores, theta=theta[:n])
        rate1, secs1 = cpu_rate(om, 2, 1, theta=theta[:2])
        rec["cpu_baseline"] = {"value": rate, "unit": UNIT, "cores": cores, "kind": "port",
                               "sample": f"{n} particles of the same model on {cores} pthreads, {secs:.1f} s; {CPU_NOTE}",
                               "one_thread": {"value": rate1, "unit": UNIT, "cores": 1}}
    return rec


def run_sweep(pkg, ctx, model, gm, P, seed, barrier, n_gen=3):
    """BASELINE C5 as a sweep: n_gen chained generations of P particles - proposals, scoring (sharded + all-gathered in
    the library), accept, epsilon, PMC weights (k_pmc_mixture, sharded), covariance - every step a user's pmc_abc does."""
    rng = np.random.default_rng(77)                       # identical on every rank
    eps = 1.0
    gens = []
    theta_prev = w_prev = tau_sq = None
    barrier()
    t_all = time.perf_counter()
    for g in range(1, n_gen + 1):
        t0 = time.perf_counter()
        if g == 1:
            theta = rng.uniform(*PRIOR, size=(P, 1))
        else:
            # PMC proposals on the device (itjl_pmc_propose): every rank draws the same matrix from the Philox stream
            theta = pkg.draw_theta_pmc_device(theta_prev, w_prev, tau_sq, P, seed=seed, generation=g, ctx=ctx)
        t1 = time.perf_counter()
        d, isacc, n_total, n_acc = gm.generation(theta, eps, P + 1, seed=seed, generation=g)
        t2 = time.perf_counter()
        acc = isacc[:n_total]
        theta_acc = theta[:n_total][acc]
        t2a = time.perf_counter()
        new_eps = pkg.select_epsilon(d[:n_total], eps, current_acc_rate=n_acc / n_total, iteration=g, total_iterations=n_gen, ctx=ctx)
        t3 = time.perf_counter()
        if g == 1:
            w = np.full(theta_acc.shape[0], 1.0 / theta_acc.shape[0])
            tsq = 2.0 * np.atleast_2d(np.cov(theta_acc, rowvar=False)) + 1e-6
        else:
            w = pkg.calc_weights(theta_prev, theta_acc, tau_sq, w_prev, model.prior, ctx=ctx)
            tsq = 2.0 * pkg.weighted_covar(theta_acc, w, ctx=ctx)
        t4 = time.perf_counter()
        gens.append({"generation": g, "epsilon": eps, "n_accepted": int(n_acc), "propose_host_ms": 1e3 * (t1 - t0),
                     "score_ms": 1e3 * (t2 - t1), "accept_epsilon_host_ms": 1e3 * (t3 - t2), "accept_compaction_ms": 1e3 * (t2a - t2), "weights_cov_ms": 1e3 * (t4 - t3),
                     "weight_pairs": 0 if g == 1 else int(theta_acc.shape[0]) * int(theta_prev.shape[0]),
                     "posterior_mean": float(np.average(theta_acc[:, 0], weights=w))})
        theta_prev, w_prev, tau_sq, eps = theta_acc, w, tsq, new_eps
    barrier()
    secs = time.perf_counter() - t_all
    return secs, gens


def run_gpu(args):
    import torch
    import torch.distributed as dist
    import __graft_entry__ as ge

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    distributed = world > 1
    torch.cuda.set_device(local_rank)
    if distributed:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    pkg = ge.load_package()
    from importlib import import_module
    idist = import_module("itjl_b200.distributed")

    ctx = pkg.default_context(local_rank)
    if distributed:
        idist.attach_library_comm(ctx)                              # the library's own NCCL communicator
    data, t = observed_data_host()
    model = pkg.one_timescale_model(data, t, "abc", prior=[pkg.Uniform(*PRIOR)])
    gm = model.gpu_model(ctx)
    flops_per_sample, samples_per_particle = gm.work()
    kernel_name, tensor_flops_per_sample = gm.kernel_info()

    P = int(args.particles)
    lo, hi = idist.shard_bounds(P, world, rank)
    n_local = hi - lo
    stream = torch.cuda.ExternalStream(ctx.stream, device=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)

    fp32_peak = ctx.fp32_peak(5)                                   # TFLOP/s, measured now

    rng = np.random.default_rng(1234)                              # same proposals on every rank
    n_sets = 2
    theta_host = [torch.from_numpy(rng.uniform(*PRIOR, size=P)).pin_memory() for _ in range(n_sets)]
    with torch.cuda.stream(stream):
        theta_dev = [th.to(dev, non_blocking=True) for th in theta_host]
        dist_all = torch.empty(P, dtype=torch.float64, device=dev)
        flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    stream.synchronize()

    def barrier():
        if distributed:
            dist.barrier()
        torch.cuda.synchronize()

    def step_resident(k):
        """inputs already in HBM; one generation scored by this rank's share, ONE in-library all-gather of distances."""
        th = theta_dev[k % n_sets]
        with torch.cuda.stream(stream):
            flush.fill_(k & 0xFF)                                  # L2 flush between steps
        gm.distances_sharded_dev(th.data_ptr(), P, 1, dist_all.data_ptr(), seed=args.seed, generation=k + 1)

    def step_e2e(k):
        """public host API: host theta -> H2D -> kernels -> in-library all-gather -> accept -> D2H distances + mask."""
        th = theta_host[k % n_sets].numpy()
        with torch.cuda.stream(stream):
            flush.fill_(k & 0xFF)
        return gm.generation(th.reshape(-1, 1), epsilon=0.01, min_accepted=P + 1, seed=args.seed, generation=k + 1)

    sweep_mode = args.config == "c5sweep"
    # ---- warm-up ------------------------------------------------------------------
    for w in range(args.warmup):
        step_resident(w)
    barrier()

    # ---- timed: K resident steps ----------------------------------------------------
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ctx.timing_reset()
    launches0 = ctx.launch_count
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record(stream)
    for k in range(args.steps):
        step_resident(args.warmup + k)
    e1.record(stream)
    barrier()
    ms = e0.elapsed_time(e1)
    n_kern, kern_ms = ctx.timing_get()
    ctx.timing_stop()
    launches = ctx.launch_count - launches0
    rank_ms = [ms]
    if distributed:
        tms = torch.tensor([ms], dtype=torch.float64, device=dev)
        every = torch.empty(world, dtype=torch.float64, device=dev)
        dist.all_gather_into_tensor(every, tms)
        rank_ms = [float(v) for v in every.tolist()]
        ms = max(rank_ms)                                          # max over ranks
    clocks = sampler.stop() if rank == 0 else None

    # ---- timed: end-to-end through the host API ----------------------------------------
    k_e2e = max(1, min(args.steps, 3))
    step_e2e(0)
    barrier()
    t0 = torch.cuda.Event(enable_timing=True)
    t1 = torch.cuda.Event(enable_timing=True)
    t0.record(stream)
    for k in range(k_e2e):
        step_e2e(args.warmup + k)
    t1.record(stream)
    barrier()
    e2e_ms = t0.elapsed_time(t1)
    if distributed:
        tms = torch.tensor([e2e_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
        e2e_ms = float(tms.item())

    # ---- the 3-generation sweep (every rank: its host steps are redundant, its device steps sharded) ---------
    sweep = None
    if not args.no_sub or sweep_mode:
        run_sweep(pkg, ctx, model, gm, P, args.seed, barrier, n_gen=2)     # warm-up: two generations at full size (every buffer the sweep grows is allocated here, not inside the timed loop)
        sweep_s, gens = run_sweep(pkg, ctx, model, gm, P, args.seed, barrier)
        if distributed:
            tms = torch.tensor([sweep_s], dtype=torch.float64, device=dev)
            dist.all_reduce(tms, op=dist.ReduceOp.MAX)
            sweep_s = float(tms.item())
        n_gen = len(gens)
        sweep = {"metric": METRIC, "unit": UNIT, "value": n_gen * P * samples_per_particle / sweep_s, "seconds": sweep_s,
                 "config": {"workload": f"C5 sweep: {n_gen} chained generations x {P} particles (prior -> PMC proposals), epsilon by "
                                        "select_epsilon, PMC weights + covariance on the device between generations",
                            "parallelism": f"particle-shard x{world}"},
                 "generations": gens,
                 "host_ms_total": sum(g["propose_host_ms"] + g["accept_epsilon_host_ms"] for g in gens),
                 "score_ms_total": sum(g["score_ms"] for g in gens),
                 "weights_cov_ms_total": sum(g["weights_cov_ms"] for g in gens),
                 "note": "wall clock of the whole host-driven loop (max over ranks); PMC proposals are drawn on the device "
                         "(k_pmc_propose, every rank the same matrix), epsilon quantiles and the accept compaction run on the host "
                         "of every rank, scoring and weights are sharded over the GPUs"}

    samples_per_step = P * samples_per_particle
    value = args.steps * samples_per_step / (ms * 1e-3)
    e2e_value = k_e2e * samples_per_step / (e2e_ms * 1e-3)
    flops_per_launch = flops_per_sample * n_local * samples_per_particle
    avg_kernel_ms = kern_ms / max(n_kern, 1)
    achieved = flops_per_launch / (avg_kernel_ms * 1e-3) / 1e12

    if rank != 0:
        if distributed:
            dist.destroy_process_group()
        return

    tensor_peak, tensor_src = measured_peak("bf16_tflops_sustained", 1400.0)
    hbm_peak, hbm_src = measured_peak("hbm_gbs", 6400.0)
    plan_terms = None
    if kernel_name == "k_abc_acf_tc":
        import ctypes
        plan = (ctypes.c_int32 * 16)()
        pkg._lib.lib().itjl_tc_plan(N_T, int(model.n_lags), N_TRIALS, 232448, ctx.sm_count, plan)
        plan_terms = int(plan[7])
        executed = tensor_flops_per_sample * n_local * samples_per_particle / (avg_kernel_ms * 1e-3) / 1e12
        roofline = {"bound": "tensor", "achieved": achieved, "peak": tensor_peak, "unit": "TFLOP/s",
                    "frac": achieved / tensor_peak,
                    # dram__bytes_read.sum + dram__bytes_write.sum of one 10^6-particle launch, ncu --set full, parsed from
                    # the committed summary (only quoted for that launch size)
                    "traffic": (ncu_traffic("r02_k_abc_acf_tc_fullsize_ncu_summary.txt") or
                                ncu_traffic("r01_k_abc_acf_tc_v8_fullsize_ncu_summary.txt")) if n_local == 1000000 else None,
                    "traffic_note": "compute-bound kernel: ~8 MB of DRAM reads (theta) per 10^6-particle launch against 2.5e11 "
                                    "samples simulated in shared memory - trajectories never leave the SM",
                    "kernel": kernel_name, "avg_launch_ms": avg_kernel_ms, "launches_timed": int(n_kern),
                    "flops_per_sample": flops_per_sample, "peak_source": tensor_src + " (kernel timed inside a long step)",
                    "executed_tensor_tflops": executed, "executed_tensor_frac": executed / tensor_peak,
                    "executed_tensor_flops_per_sample": tensor_flops_per_sample,
                    "fp32_peak_measured": fp32_peak, "achieved_over_fp32_peak": achieved / fp32_peak if fp32_peak else None,
                    "product_terms": plan_terms,
                    "note": "algorithmic flops (2 n_lags + 8 per sample) exceed the FP32-FMA roofline; x' is fed to tcgen05 as fp16 "
                            "(the error model keeps hi*hi products only at this shape, hi/lo pairs with 2 or 3 MMAs per product for "
                            "smaller particles) on padded tiles, so the executed tensor flops are a multiple of the algorithmic "
                            "count (both reported); the simulate phase (Philox + scan, CUDA cores) runs under the MMAs"}
        dtype = ("f16 x f16 -> f32 tensor-core lag sums (hi*hi term only at this shape), fp32 simulation, fp64 distances"
                 if plan_terms == 1 else f"f16 hi/lo x f16 hi/lo -> f32 ({plan_terms} product terms), fp32 simulation, fp64 distances")
    else:
        roofline = {"bound": "fp32", "achieved": achieved, "peak": fp32_peak, "unit": "TFLOP/s",
                    "frac": achieved / fp32_peak if fp32_peak else None, "traffic": None,
                    "kernel": kernel_name, "avg_launch_ms": avg_kernel_ms, "launches_timed": int(n_kern),
                    "flops_per_sample": flops_per_sample,
                    "peak_source": "FP32 FMA issue-rate microbenchmark (itjl_fp32_peak) measured in this run; "
                                   "MEASURED_PEAKS.json has no FP32 entry; nominal 148 SM x 128 x 2 x 1.965 GHz = 74.4"}
        dtype = "f32 (FFMA lag sums), fp64 distances"

    sub = None
    cpu = None
    want_cpu = not args.no_cpu_baseline
    if world == 1 and not args.no_sub:
        sub = {}
        # the same workload with all three product terms (what smaller particles get): the price of the exact split
        with ctx.tuned(tc_terms=3):
            g3 = pkg.one_timescale_model(data, t, "abc", prior=[pkg.Uniform(*PRIOR)]).gpu_model(ctx)
        th3 = np.random.default_rng(2).uniform(*PRIOR, size=(100000, 1))
        ms3, _ = time_model(pkg, ctx, torch, stream, g3, th3, 5)
        sub["c5_three_product_terms"] = {"metric": METRIC, "unit": UNIT, "value": 100000 * samples_per_particle / (ms3 * 1e-3),
                                         "config": {"workload": "C5 shape, tuning.tc_terms = 3 (hi*hi + hi*lo + lo*hi), 10^5 particles"}}
        sub["c2_acw"] = sub_c2(pkg, ctx, torch, stream, hbm_peak, hbm_src, want_cpu)
        d3, t3 = observed_c3()
        m3 = pkg.one_timescale_with_missing_model(d3, t3, "abc", prior=[pkg.Uniform(0.0, 5.0)])

        def om3():
            from oracle import models as omodels
            return omodels.one_timescale_with_missing_model(d3, t3, prior=[omodels.Uniform(0.0, 5.0)])
        sub["c3_missing"] = sub_model(pkg, ctx, torch, stream, "C3 one_timescale_with_missing_model, 100 trials x 5000, 10 % NaN gaps",
                                      m3, om3, np.random.default_rng(3).uniform(0.05, 5.0, size=(20000, 1)), fp32_peak,
                                      tensor_peak, tensor_src, want_cpu)
        d4, t4 = observed_c4()
        pri4 = [pkg.Normal(.1, .1), pkg.Normal(10, 5), pkg.Uniform(0, 1)]
        m4 = pkg.one_timescale_and_osc_model(d4, t4, "abc", prior=pri4)

        def om4():
            from oracle import models as omodels
            return omodels.one_timescale_and_osc_model(d4, t4, prior=[omodels.Normal(.1, .1), omodels.Normal(10, 5), omodels.Uniform(0, 1)])
        r4 = np.random.default_rng(4)
        n4 = 14 * ctx.sm_count                      # one CTA per particle, one CTA per SM: whole waves (2072 on a B200)
        th4 = np.stack([np.abs(r4.normal(.1, .1, n4)) + 0.005, r4.normal(10, 5, n4), r4.uniform(0, 1, n4)], axis=1)
        sub["c4_osc_psd"] = sub_model(pkg, ctx, torch, stream, "C4 one_timescale_and_osc_model, PSD summary, 100 trials x 10000",
                                      m4, om4, th4, fp32_peak, tensor_peak, tensor_src, want_cpu)
        # f1: one_timescale_model(summary_method = :psd), DSP periodogram (nfft = 5000), 20 trials x 5000
        d1, t1 = observed_c3()
        d1 = np.nan_to_num(d1[:20])                                  # complete data (the C3 gaps filled with zeros)
        mp = pkg.one_timescale_model(d1, t1, "abc", summary_method="psd", prior=[pkg.Uniform(0.0, 5.0)])

        def omp():
            from oracle import models as omodels
            return omodels.one_timescale_model(d1, t1, summary_method="psd", prior=[omodels.Uniform(0.0, 5.0)])
        sub["f1_psd_periodogram"] = sub_model(pkg, ctx, torch, stream,
                                              "one_timescale_model(:psd): DSP periodogram summary, 20 trials x 5000, nfft = 5000",
                                              mp, omp, np.random.default_rng(6).uniform(0.05, 5.0, size=(20000, 1)), fp32_peak,
                                              tensor_peak, tensor_src, want_cpu)
        sub["f3_acw_knee"] = sub_knee(pkg, ctx, want_cpu)
    if sweep is not None:
        sub = sub or {}
        sub["c5_sweep"] = sweep
    if world == 1 and want_cpu:
        cpu = cpu_baseline()

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": dtype, "data": "synthetic",
            "config": {"workload": "C5 aABC generation: 10^6 particles x 50 trials x 5000 samples, OU tau~U(0.05,5), "
                                   "trial-mean ACF + linear distance", "particles_per_step": P,
                       "n_trials": N_TRIALS, "n_t": N_T, "n_lags": int(model.n_lags), "parallelism": f"particle-shard x{world}",
                       "l2": "256 MiB buffer written between steps (inputs are 8 MB/step; the kernel is compute bound)",
                       "collective": "one in-library ncclAllGather of per-particle distances per step (8 B x particles, device "
                                     "to device)" if distributed else "none"},
            "roofline": roofline,
            "cpu_baseline": cpu,
            "e2e": {"value": e2e_value, "unit": UNIT, "steps": k_e2e,
                    "h2d_bytes_per_step": int(8 * P), "d2h_bytes_per_step": int(9 * P)},
            "gpu_launches": int(launches),
            "rank_ms_per_step": [v / args.steps for v in rank_ms],
            "clocks": clocks,
            "sub": sub}
    if sweep_mode and sweep is not None:
        # --config c5sweep: the headline value is the whole 3-generation sweep; the generation-scoring numbers move to `sub`
        line["sub"] = dict(sub or {}, c5_generation={"value": value, "unit": UNIT, "ms_per_step": ms / args.steps})
        line.update(value=sweep["value"], ms_per_step=1e3 * sweep["seconds"], steps=1,
                    config=dict(line["config"], workload=sweep["config"]["workload"]))
    print(json.dumps(line), flush=True)
    if distributed:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="c5", choices=["c5", "c5sweep"])
    ap.add_argument("--particles", type=int, default=1_000_000)
    ap.add_argument("--seed", type=int, default=2026)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-sub", action="store_true", help="skip the C2 / C3 / C4 / sweep sub-records")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
