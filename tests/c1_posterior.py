#!/usr/bin/env python
"""BASELINE config C1 (the README example, /root/reference/README.md:43-55) through the PMC-ABC
sampler of both arms, many seeds:

    data  = generate_ou_process(0.3, 1.0, 1/500, 10.0, 2)              (2 trials x 5000)
    model = one_timescale_model(data, time, :abc)                       (informed prior Normal(tau_hat, 20))
    int_fit(model)                                                      (get_param_dict_abc defaults)

and the docs variant (docs/src/practice/practice_5_bayesian.md: 30 trials, prior Uniform(0, 5)... run
here with the same 2-trial data and `--variant uniform`).

    python tests/c1_posterior.py --arm oracle --seeds 24 --out tests/golden/c1_posterior_oracle.json
    python tests/c1_posterior.py --arm gpu    --seeds 24 --out gpurun_out/c1_posterior_gpu.json

The oracle arm is the CPU restatement of the reference's serial sampler (oracle/abc.py over the C loop body);
it is test infrastructure: its output is committed as a fixture and `tests/test_gpu_c1_posterior.py` holds
the GPU arm's posterior mean / std distribution against it (north_star correctness item 3).  Both arms share
the observed data (generated once by the oracle's exact-OU generator, seed 123) and nothing else: the
sampler streams differ (NumPy generator seeds 1000 + s vs 2000 + s), as two runs of the reference with
different `rng`s would.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

DT, N_T, N_TRIALS, TAU = 0.002, 5000, 2, 0.3


def observed():
    from oracle import ou as oou
    data = oou.generate_ou_process(TAU, 1.0, DT, N_T * DT, N_TRIALS, seed=123, n_t=N_T)
    return data, DT * np.arange(1, N_T + 1)


def summarise(theta, weights, eps_hist, acc_hist, n_steps, secs):
    th = np.asarray(theta, dtype=np.float64).reshape(-1)
    w = np.asarray(weights, dtype=np.float64).reshape(-1)
    ok = th.size > 0 and w.size == th.size and w.sum() > 0
    return {"n": int(th.size), "mean": float(th.mean()) if th.size else None,
            "std": float(th.std()) if th.size else None,
            "wmean": float(np.average(th, weights=w)) if ok else None,
            "median": float(np.median(th)) if th.size else None,
            "q05": float(np.percentile(th, 5)) if th.size else None,
            "q95": float(np.percentile(th, 95)) if th.size else None,
            "frac_below_1": float((th < 1.0).mean()) if th.size else None,
            "steps": int(n_steps), "epsilon_last": float(eps_hist[-1]), "epsilon_first": float(eps_hist[0]),
            "acc_rate_last": float(acc_hist[-1]), "seconds": secs}


def run_oracle(variant, seeds, threads):
    from oracle import abc as oabc
    from oracle import cport
    from oracle import models as omodels
    data, t = observed()
    prior = None if variant == "informed" else [omodels.Uniform(0.0, 5.0)]
    om = omodels.one_timescale_model(data, t, prior=prior)
    params = oabc.get_param_dict_abc()
    out = []
    for s in seeds:
        rng = np.random.default_rng(1000 + s)

        def batch_fn(thetas, first, generation, s=s):
            return cport.abc_distances(om, thetas, seed=s, generation=generation, particle_offset=first,
                                       n_threads=threads)
        t0 = time.perf_counter()
        rec = oabc.pmc_abc(om.prior, None, rng, distance_batch_fn=batch_fn, **params)
        r = summarise(rec[-1]["theta_accepted"], rec[-1]["weights"], [x["epsilon"] for x in rec],
                      [x["n_accepted"] / x["n_total"] for x in rec], len(rec), time.perf_counter() - t0)
        r["seed"] = int(s)
        out.append(r)
        print(json.dumps(r), file=sys.stderr, flush=True)
    return om, out


def run_gpu(variant, seeds):
    import __graft_entry__ as ge
    pkg = ge.load_package()
    data, t = observed()
    prior = None if variant == "informed" else [pkg.Uniform(0.0, 5.0)]
    gm = pkg.one_timescale_model(data, t, "abc", prior=prior)
    out = []
    for s in seeds:
        t0 = time.perf_counter()
        res = pkg.int_fit(gm, dict(show_progress=False, verbose=False), rng=np.random.default_rng(2000 + s), seed=s)
        r = summarise(res.final_theta, res.final_weights, res.epsilon_history, res.acc_rate_history,
                      len(res.epsilon_history), time.perf_counter() - t0)
        r["seed"] = int(s)
        out.append(r)
    return gm, out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--arm", choices=["oracle", "gpu"], required=True)
    ap.add_argument("--variant", choices=["informed", "uniform"], default="informed")
    ap.add_argument("--seeds", type=int, default=24)
    ap.add_argument("--threads", type=int, default=0)
    ap.add_argument("--out", required=True)
    a = ap.parse_args()
    seeds = list(range(a.seeds))
    m, runs = (run_oracle(a.variant, seeds, a.threads) if a.arm == "oracle" else run_gpu(a.variant, seeds))
    pri = m.prior[0]
    doc = {"config": "C1 README example: OU tau=0.3, 2 trials x 5000 samples, dt=0.002; one_timescale_model(:abc); "
                     "int_fit with get_param_dict_abc defaults",
           "arm": a.arm, "variant": a.variant, "n_lags": int(m.n_lags), "prior": repr(pri), "runs": runs}
    os.makedirs(os.path.dirname(os.path.abspath(a.out)), exist_ok=True)
    with open(a.out, "w") as f:
        json.dump(doc, f, indent=1)
    means = np.array([r["mean"] for r in runs if r["mean"] is not None])
    stds = np.array([r["std"] for r in runs if r["std"] is not None])
    print(json.dumps({"arm": a.arm, "variant": a.variant, "n_runs": len(runs), "posterior_mean_avg": float(means.mean()),
                      "posterior_mean_sd_over_seeds": float(means.std(ddof=1)) if means.size > 1 else None,
                      "posterior_std_avg": float(stds.mean())}))


if __name__ == "__main__":
    main()
