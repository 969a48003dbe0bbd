"""world_size = 2 (and 3) gloo runs of the particle-sharding logic on CPU: the sharded
scorer must return exactly what a single rank computes, for any chunking."""
import os
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _distance(theta, generation, offset):
    # deterministic function of (theta, global particle index, generation): stands in for
    # the fused kernel, whose Philox streams are keyed on the same triple
    idx = offset + np.arange(theta.shape[0])
    return (theta[:, 0] - 1.0) ** 2 + 1e-3 * ((idx * 2654435761 + generation * 97) % 1000) / 1000.0


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    import __graft_entry__ as ge
    pkg = ge.load_package()
    from importlib import import_module
    distributed = import_module("itjl_b200.distributed")
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        scorer = distributed.make_sharded_scorer(_distance)

        class M:
            prior = [pkg.Uniform(0, 3)]
            fit_method = "abc"
        rng = np.random.default_rng(0)
        props = rng.random((1500, 1)) * 3
        res = pkg.basic_abc(M(), 0.02, 1500, 37, scorer=scorer, proposals=props, generation=4)
        from oracle import abc as oabc          # weight / covariance steps injected (device kernels in the product)
        res2 = pkg.pmc_abc(M(), epsilon_0=0.5, max_iter=800, min_accepted=50, steps=3,
                           rng=np.random.default_rng(5), scorer=scorer, target_epsilon=1e-9,
                           calc_weights_fn=oabc.calc_weights, weighted_covar_fn=oabc.weighted_covar)
        out[rank] = (res["n_total"], res["n_accepted"], res["distances"], res["theta_accepted"],
                     res2.final_theta, res2.epsilon_history)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_generation_matches_single_rank(pkg, world):
    port = 29500 + os.getpid() % 2000 + world
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    # single-rank reference
    from importlib import import_module
    distributed = import_module("itjl_b200.distributed")

    def single(theta, epsilon, needed, generation, offset):
        d = _distance(theta, generation, offset)
        acc, n_total, n_acc = distributed.stop_and_accept(d, epsilon, needed)
        return d, acc, n_total, n_acc

    class M:
        prior = [pkg.Uniform(0, 3)]
        fit_method = "abc"
    props = np.random.default_rng(0).random((1500, 1)) * 3
    ref = pkg.basic_abc(M(), 0.02, 1500, 37, scorer=single, proposals=props, generation=4)
    from oracle import abc as oabc
    ref2 = pkg.pmc_abc(M(), epsilon_0=0.5, max_iter=800, min_accepted=50, steps=3,
                       rng=np.random.default_rng(5), scorer=single, target_epsilon=1e-9,
                       calc_weights_fn=oabc.calc_weights, weighted_covar_fn=oabc.weighted_covar)
    for r in range(world):
        n_total, n_acc, d, th, final_theta, eps_hist = out[r]
        assert (n_total, n_acc) == (ref["n_total"], ref["n_accepted"])
        assert np.array_equal(d, ref["distances"]) and np.array_equal(th, ref["theta_accepted"])
        assert np.array_equal(final_theta, ref2.final_theta) and eps_hist == ref2.epsilon_history


def test_shard_bounds_cover_everything(pkg):
    from importlib import import_module
    distributed = import_module("itjl_b200.distributed")
    for n in [0, 1, 7, 8, 1000003]:
        for w in [1, 2, 3, 8]:
            b = [distributed.shard_bounds(n, w, r) for r in range(w)]
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(w - 1))
            sizes = [hi - lo for lo, hi in b]
            assert max(sizes) - min(sizes) <= 1
