"""Chain-parallel sharding across the GPUs of one box.

Chains, (g, m) parameter points and sources are independent units: rank r of P owns a contiguous block of
chains and runs its own batched solves with no data-path collective.  The only collective is the final
measurement reduction (per-observable sums and sums of squares), a handful of doubles per parameter point.
Works with any torch.distributed backend: "nccl" on the GPUs, "gloo" in the CPU tests.
"""
from __future__ import annotations

import numpy as np


def chain_range(rank: int, world: int, nchains_total: int):
    """Block distribution: the first (nchains_total % world) ranks own one extra chain.
    Returns (first_chain, nchains_local)."""
    if not (0 <= rank < world) or nchains_total < 0:
        raise ValueError("bad rank/world/nchains")
    base, extra = divmod(nchains_total, world)
    n = base + (1 if rank < extra else 0)
    first = rank * base + min(rank, extra)
    return first, n


def owner_of(chain: int, world: int, nchains_total: int) -> int:
    base, extra = divmod(nchains_total, world)
    cut = extra * (base + 1)
    if chain < cut:
        return chain // (base + 1)
    return extra + (chain - cut) // base


def deal_by_cost(costs, world: int):
    """Cost-balanced dealing of independent units (chains of a coupling/mass scan: the cost of a chain is its expected
    CG iteration count, e.g. that of the previous solve, which grows like 1/m).  Longest-processing-time rule: units in
    descending cost order, each to the rank with the least work so far (ties to the lowest rank).  Returns one index
    array per rank, each in descending cost order -- which is also the order a rank should RUN them in: the on-chip
    solvers start chains in index order, so the long solves start first and the tail of a batch is made of short ones.
    Deterministic; every rank computes the same deal from the same costs."""
    costs = np.asarray(costs, dtype=np.float64)
    if world < 1 or costs.ndim != 1:
        raise ValueError("bad world/costs")
    order = np.argsort(-costs, kind="stable")
    load = np.zeros(world)
    mine = [[] for _ in range(world)]
    for i in order:
        r = int(np.argmin(load))
        mine[r].append(int(i))
        load[r] += costs[i]
    return [np.asarray(m, dtype=np.int64) for m in mine]


def reduce_observables(local_values, group_ids=None, ngroups=1, dist=None, device=None):
    """Final measurement reduction.  local_values: (nlocal, nobs) per-chain observables of this rank;
    group_ids: (nlocal,) parameter-point index of each local chain.  Returns per group
    (count, mean, standard error) over ALL ranks; one all_reduce of ngroups*(1+2*nobs) doubles."""
    import torch

    v = np.atleast_2d(np.asarray(local_values, dtype=np.float64))
    nlocal, nobs = v.shape if v.size else (0, v.shape[1] if v.ndim == 2 else 1)
    gid = np.zeros(nlocal, dtype=np.int64) if group_ids is None else np.asarray(group_ids, dtype=np.int64)
    acc = np.zeros((ngroups, 1 + 2 * nobs))
    for i in range(nlocal):
        acc[gid[i], 0] += 1.0
        acc[gid[i], 1:1 + nobs] += v[i]
        acc[gid[i], 1 + nobs:] += v[i] ** 2
    t = torch.from_numpy(acc)
    if device is not None:
        t = t.to(device)
    if dist is not None and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    acc = t.cpu().numpy()
    cnt = acc[:, 0]
    safe = np.maximum(cnt, 1.0)
    mean = acc[:, 1:1 + nobs] / safe[:, None]
    var = np.maximum(acc[:, 1 + nobs:] / safe[:, None] - mean ** 2, 0.0)
    err = np.sqrt(var / np.maximum(cnt - 1.0, 1.0)[:, None])
    return cnt, mean, err
