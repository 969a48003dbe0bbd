"""north_star (3) on BASELINE config C1, exactly as the README runs it (/root/reference/README.md:43-55):
2 trials x 5000 samples, one_timescale_model(data, time, :abc) with the informed prior Normal(tau_hat, 20),
int_fit with the get_param_dict_abc defaults - and the docs' Uniform(0, 5) prior on the same data.

The CPU side is a committed fixture: tests/golden/c1_posterior_oracle_{informed,uniform}.json, 24 seeds of the
oracle's restatement of the reference's serial PMC-ABC sampler over the fp64 C loop body, written by
`python tests/c1_posterior.py --arm oracle ...` (3 minutes of CPU each).  Here the GPU arm runs the same 24
seeds and both posterior summaries (mean and std of final_theta) must agree within Monte-Carlo error.

What the informed-prior run looks like in BOTH arms (the round-1 question "posterior 11.4 +- 13.5 for a true tau
of 0.3?"): Normal(0.2, 20) proposes |tau| of tens of seconds against 10 s of data; for tau >> 1 s the simulated
ACF is flat over the 297 lags (0.6 s) that are scored, so the distance sits on a plateau at ~0.2 whatever tau
is, and select_epsilon's floor `max(q25, eps (1 - alpha))` (abc.jl:746-752) cannot go below the 25 % quantile
of a population that mostly lives on that plateau: epsilon stalls at 0.19-0.20 for all 30 steps and the
'posterior' is the prior's positive half reshaped by the PMC kernel (mean 13.1, std 12.9 in the oracle).
That is the reference algorithm's behaviour under its default prior, not a defect of the negative-tau path;
with Uniform(0, 5) the same sampler recovers tau = 0.276 +- 0.076.
"""
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
N_SEEDS = 24


def _gpu_runs(variant):
    import c1_posterior as c1
    _, runs = c1.run_gpu(variant, list(range(N_SEEDS)))
    return runs


def _fixture(variant):
    with open(os.path.join(HERE, "golden", f"c1_posterior_oracle_{variant}.json")) as f:
        return json.load(f)


@pytest.mark.parametrize("variant", ["informed", "uniform"])
def test_c1_posterior_mean_and_std_within_mc_error_of_the_oracle(pkg, ctx, variant):
    import sys
    sys.path.insert(0, HERE)
    fx = _fixture(variant)
    runs = _gpu_runs(variant)
    assert len(fx["runs"]) == N_SEEDS == len(runs)
    out = {"variant": variant, "n_seeds": N_SEEDS}
    for key in ("mean", "std"):
        o = np.array([r[key] for r in fx["runs"]])
        g = np.array([r[key] for r in runs])
        se = np.sqrt(o.var(ddof=1) / o.size + g.var(ddof=1) / g.size)        # MC error of the difference of the two averages
        z = (g.mean() - o.mean()) / se
        out[key] = {"oracle": float(o.mean()), "gpu": float(g.mean()), "se": float(se), "z": float(z),
                    "oracle_sd_over_seeds": float(o.std(ddof=1)), "gpu_sd_over_seeds": float(g.std(ddof=1))}
        assert abs(z) < 3.5, (variant, key, out[key])
        # seed-to-seed scatter of the summary itself is the same in both arms (F-test-like band for 24 vs 24)
        assert 0.45 < g.std(ddof=1) / o.std(ddof=1) < 2.2, (variant, key, out[key])
    eo = np.array([r["epsilon_last"] for r in fx["runs"]])
    eg = np.array([r["epsilon_last"] for r in runs])
    out["epsilon_last"] = {"oracle": float(eo.mean()), "gpu": float(eg.mean())}
    assert abs(eg.mean() - eo.mean()) < 3.5 * np.sqrt(eo.var(ddof=1) / eo.size + eg.var(ddof=1) / eg.size) + 1e-12
    assert [r["steps"] for r in runs] == [r["steps"] for r in fx["runs"]] or variant == "uniform"
    if variant == "uniform":
        # recovers the generating timescale as far as TWO 10 s trials determine it: the posterior follows the data's own
        # ACF (this realisation of the Philox4x32-7 stream: oracle 0.394, n_lags 518; the 10-round stream's: 0.276)
        assert abs(out["mean"]["gpu"] - 0.3) < 0.15 and abs(out["mean"]["gpu"] - out["mean"]["oracle"]) < 0.02
    out["gpu_seconds_total"] = float(sum(r["seconds"] for r in runs))
    print("C1 posterior", json.dumps(out))
    os.makedirs(os.path.join(os.path.dirname(HERE), "gpurun_out"), exist_ok=True)
    with open(os.path.join(os.path.dirname(HERE), "gpurun_out", f"c1_posterior_{variant}.json"), "w") as f:
        json.dump({"summary": out, "gpu_runs": runs}, f, indent=1)
