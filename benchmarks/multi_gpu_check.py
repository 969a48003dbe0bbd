#!/usr/bin/env python
"""torchrun --nproc-per-node N benchmarks/multi_gpu_check.py: the particle-sharded scorer over N GPUs (NCCL
all-gather of distances) must return exactly what one GPU returns - distances bit-identical, same stop index,
same accepted set, same PMC history (results invariant to the GPU count)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist
import __graft_entry__ as ge

rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
torch.cuda.set_device(local)
os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
pkg = ge.load_package()


def _tun(ctx=None):
    """ITJL_* environment switches of this script -> the context's itjl_tuning (benchmarks/_tuning.py)."""
    import os as _os, sys as _sys
    _sys.path.insert(0, _os.path.dirname(_os.path.abspath(__file__)))
    import _tuning
    return _tuning.from_env(ctx or pkg.default_context())

from importlib import import_module
distributed = import_module("itjl_b200.distributed")
from oracle import ou as oou
ctx = pkg.default_context(local)
ok = True
for n_t, n_trials, n_lags in [(5000, 10, None), (3000, 4, 700)]:
    dt = 0.002
    data = oou.generate_ou_process(0.3, 1.0, dt, n_t * dt, n_trials, seed=5, n_t=n_t)
    t = dt * np.arange(1, n_t + 1)
    model = pkg.one_timescale_model(data, t, "abc", prior=[pkg.Uniform(0.05, 5.0)], n_lags=n_lags)
    scorer = distributed.gpu_sharded_scorer(model, ctx=ctx, seed=3)
    props = np.random.default_rng(1).uniform(0.05, 5.0, size=(4001, 1))
    d_all = model.gpu_model(_tun(ctx)).distances(props, seed=3, generation=2)           # every rank: the whole batch on its own GPU
    eps = float(np.sort(d_all)[400] * 0.5 + np.sort(d_all)[401] * 0.5)
    sharded = pkg.basic_abc(model, eps, 4001, 333, scorer=scorer, proposals=props, generation=2)
    single = pkg.basic_abc(model, eps, 4001, 333, proposals=props, seed=3, generation=2, ctx=ctx)
    same = (sharded["n_total"] == single["n_total"] and sharded["n_accepted"] == single["n_accepted"]
            and np.array_equal(sharded["distances"], single["distances"]) and np.array_equal(sharded["theta_accepted"], single["theta_accepted"])
            and np.array_equal(sharded["distances"], d_all[: sharded["n_total"]]))
    r_sh = pkg.pmc_abc(model, epsilon_0=1.0, max_iter=3000, min_accepted=100, steps=4, rng=np.random.default_rng(7), scorer=scorer, target_epsilon=1e-9)
    r_si = pkg.pmc_abc(model, epsilon_0=1.0, max_iter=3000, min_accepted=100, steps=4, rng=np.random.default_rng(7), seed=3, ctx=ctx, target_epsilon=1e-9)
    same_pmc = np.array_equal(r_sh.final_theta, r_si.final_theta) and r_sh.epsilon_history == r_si.epsilon_history
    print(f"rank {rank}/{world}: n_t={n_t} n_lags={model.n_lags}: basic_abc identical={same} (n_total {sharded['n_total']}, accepted {sharded['n_accepted']}); "
          f"pmc_abc identical={same_pmc} (eps history {[round(e, 5) for e in r_sh.epsilon_history]})", flush=True)
    ok = ok and same and same_pmc
flag = torch.tensor([0 if ok else 1], device=torch.device("cuda", local))
dist.all_reduce(flag)
dist.destroy_process_group()
if rank == 0:
    print("MULTI-GPU CHECK", "PASSED" if flag.item() == 0 else "FAILED")
sys.exit(0 if flag.item() == 0 else 1)
