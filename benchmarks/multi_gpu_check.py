#!/usr/bin/env python
"""torchrun --nproc-per-node N benchmarks/multi_gpu_check.py: particles sharded over N GPUs INSIDE the library
(itjl_ctx_attach_comm: one process per GPU, ncclAllGather of distances / weights device to device) must return
exactly what one GPU returns - distances bit-identical, same stop index, same accepted set, same PMC history,
same PMC weights (results invariant to the GPU count).  Rank 0 then repeats the comparison with ONE
process driving all N GPUs (itjl_ctx_create_multi, what a Julia host uses)."""
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
from importlib import import_module
distributed = import_module("itjl_b200.distributed")
from oracle import ou as oou

ctx = pkg.Context(local)                     # joins the communicator: its calls are collective
distributed.attach_library_comm(ctx)
solo = pkg.Context(local)                    # plain single-GPU context of this rank
assert ctx.world() == (rank, world, 1) and solo.world() == (0, 1, 1)


def compare(model, sharded_ctx, tag):
    scorer = distributed.gpu_sharded_scorer(model, ctx=sharded_ctx, seed=3) if sharded_ctx is ctx else None
    props = np.random.default_rng(1).uniform(0.05, 5.0, size=(4001, 1))
    d_all = model.gpu_model(solo).distances(props, seed=3, generation=2)
    d_sh = model.gpu_model(sharded_ctx).distances(props, seed=3, generation=2)
    eps = float(np.sort(d_all)[400] * 0.5 + np.sort(d_all)[401] * 0.5)
    kw = dict(scorer=scorer) if scorer else dict(ctx=sharded_ctx, seed=3)
    sharded = pkg.basic_abc(model, eps, 4001, 333, proposals=props, generation=2, **kw)
    single = pkg.basic_abc(model, eps, 4001, 333, proposals=props, seed=3, generation=2, ctx=solo)
    same = (sharded["n_total"] == single["n_total"] and sharded["n_accepted"] == single["n_accepted"]
            and np.array_equal(sharded["distances"], single["distances"]) and np.array_equal(sharded["theta_accepted"], single["theta_accepted"])
            and np.array_equal(d_sh, d_all))
    pk = dict(epsilon_0=1.0, max_iter=3000, min_accepted=100, steps=4, target_epsilon=1e-9)
    r_sh = pkg.pmc_abc(model, rng=np.random.default_rng(7), seed=3, ctx=sharded_ctx, **pk)
    r_si = pkg.pmc_abc(model, rng=np.random.default_rng(7), seed=3, ctx=solo, **pk)
    same_pmc = (np.array_equal(r_sh.final_theta, r_si.final_theta) and r_sh.epsilon_history == r_si.epsilon_history
                and np.array_equal(r_sh.final_weights, r_si.final_weights))
    # PMC weights at a size where the rows really spread over the GPUs (fp32 pair sums above 1e8 pairs)
    rg = np.random.default_rng(3)
    th_prev, th = rg.gamma(4.0, 0.1, (15000, 1)), rg.gamma(4.0, 0.1, (9001, 1))
    w = rg.uniform(0.5, 1.5, 15000); w /= w.sum()
    cov = np.array([[2.0 * th_prev.var()]])
    w_sh = pkg.calc_weights(th_prev, th, cov, w, [pkg.Uniform(0, 5)], ctx=sharded_ctx)
    w_si = pkg.calc_weights(th_prev, th, cov, w, [pkg.Uniform(0, 5)], ctx=solo)
    same_w = np.array_equal(w_sh, w_si)
    # the one-parameter grid path (>= 32768 new particles): every rank builds the same density grid and interpolates
    # its own rows
    th_big = rg.gamma(4.0, 0.1, (40001, 1))
    g_sh = pkg.calc_weights(th_prev, th_big, cov, w, [pkg.Uniform(0, 5)], ctx=sharded_ctx)
    g_si = pkg.calc_weights(th_prev, th_big, cov, w, [pkg.Uniform(0, 5)], ctx=solo)
    same_w = same_w and np.array_equal(g_sh, g_si)
    print(f"[{tag}] rank {rank}/{world}: n_lags={model.n_lags}: basic_abc identical={same} (n_total {sharded['n_total']}, accepted "
          f"{sharded['n_accepted']}); pmc_abc identical={same_pmc} (eps history {[round(e, 5) for e in r_sh.epsilon_history]}); "
          f"calc_weights identical={same_w}", flush=True)
    return same and same_pmc and same_w


ok = True
models = []
for n_t, n_trials, n_lags in [(5000, 10, None), (3000, 4, 700)]:
    dt = 0.002
    data = oou.generate_ou_process(0.3, 1.0, dt, n_t * dt, n_trials, seed=5, n_t=n_t)
    t = dt * np.arange(1, n_t + 1)
    model = pkg.one_timescale_model(data, t, "abc", prior=[pkg.Uniform(0.05, 5.0)], n_lags=n_lags, ctx=solo)
    models.append(model)
    ok = compare(model, ctx, "one process per GPU") and ok
flag = torch.tensor([0 if ok else 1], device=torch.device("cuda", local))
dist.all_reduce(flag)
ctx.close()
dist.barrier()
if rank == 0:
    multi = pkg.Context(devices=list(range(world)))          # one process, all GPUs (ncclCommInitAll)
    assert multi.world() == (0, world, world)
    ok1 = all(compare(m, multi, "one process, all GPUs") for m in models)
    multi.close()
    flag += 0 if ok1 else 1
dist.barrier()
dist.destroy_process_group()
if rank == 0:
    print("MULTI-GPU CHECK", "PASSED" if flag.item() == 0 else "FAILED")
sys.exit(0 if flag.item() == 0 else 1)
