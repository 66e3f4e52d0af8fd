"""Host logic of the chain-parallel multi-GPU mode, covered with world_size-2 gloo processes on CPU."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from thirring2d_b200.shard import chain_range, deal_by_cost, owner_of, reduce_observables


@pytest.mark.parametrize("world,total", [(1, 256), (2, 256), (8, 2048), (3, 10), (4, 3), (8, 64)])
def test_chain_ranges_partition_the_chains(world, total):
    seen = []
    for r in range(world):
        first, n = chain_range(r, world, total)
        seen.extend(range(first, first + n))
        for c in range(first, first + n):
            assert owner_of(c, world, total) == r
    assert seen == list(range(total))
    sizes = [chain_range(r, world, total)[1] for r in range(world)]
    assert max(sizes) - min(sizes) <= 1


@pytest.mark.parametrize("world", [1, 2, 4, 8, 3])
def test_deal_by_cost_balances_a_mass_scan(world):
    """BASELINE config 5: 32 (g, m) points x 64 chains; the cost of a chain goes like 1/m.  Every chain is dealt once,
    no rank carries more than the average plus one chain's cost, and every rank runs its long chains first."""
    ms = np.tile(np.repeat([0.01, 0.03, 0.1, 0.3], 64), 8)
    costs = 30.0 / ms
    deal = deal_by_cost(costs, world)
    assert sorted(np.concatenate(deal).tolist()) == list(range(costs.size))
    loads = np.array([costs[d].sum() for d in deal])
    assert loads.max() <= costs.sum() / world + costs.max()
    assert loads.max() / loads.mean() < 1.01
    for d in deal:
        assert np.all(np.diff(costs[d]) <= 0)
    # for contrast: the same chains in mass-major order dealt in contiguous blocks (chain_range) leave one rank with
    # far more than its share
    if world > 1:
        by_mass = np.sort(costs)[::-1]
        blocks = [by_mass[f:f + n].sum() for f, n in (chain_range(r, world, costs.size) for r in range(world))]
        assert max(blocks) / (costs.sum() / world) > 1.5


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, total, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    first, n = chain_range(rank, world, total)
    # per-chain "observables" depend only on the global chain index, like bench.py's synthetic inputs
    chains = np.arange(first, first + n)
    vals = np.stack([np.sin(chains), chains.astype(float) ** 2], axis=1)
    groups = chains % 4
    cnt, mean, err = reduce_observables(vals, groups, ngroups=4, dist=dist)
    # rank-max timing reduction, as bench.py does
    t = torch.tensor([float(rank + 1)], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        q.put((cnt, mean, err, t.item()))
    dist.barrier()
    dist.destroy_process_group()


def test_world_size_2_reduction_matches_single_process():
    total, world = 37, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, total, q)) for r in range(world)]
    for p in procs:
        p.start()
    cnt, mean, err, tmax = q.get(timeout=120)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    chains = np.arange(total)
    vals = np.stack([np.sin(chains), chains.astype(float) ** 2], axis=1)
    cnt1, mean1, err1 = reduce_observables(vals, chains % 4, ngroups=4)
    assert np.array_equal(cnt, cnt1)
    assert np.allclose(mean, mean1, rtol=1e-14, atol=0) and np.allclose(err, err1, rtol=1e-10, atol=1e-12)
    assert tmax == 2.0
