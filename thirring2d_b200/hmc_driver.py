"""Batched stand-in for the reference's `hmc` program (hmc.c:847-943): same six-line `parameter` input on stdin,
same stdout keywords, many chains at once on the GPU.

    python -m thirring2d_b200.hmc_driver --nt 64 --nx 64 --chains 256 --mode adjoint --nsteps 40 < parameter

Every line the reference prints per trajectory (hmc.c:701,735,739,743,839-840) is printed per chain with a
"[chain k] " prefix, so the reference's grep-style analysis keeps working (`grep 'chain 17\\] Phase'`).
Random numbers come from the device Philox stream keyed by (seed, chain): chains are statistically, not
bit-wise, equivalent to reference runs with different seeds (use the interposed launcher for bit parity).
"""
from __future__ import annotations

import argparse
import sys

import numpy as np


def fmt_g(x: float) -> str:
    """printf("%g"): 6 significant digits, as every number the reference prints."""
    return "%g" % x


def read_parameters(stream):
    """The six scanf values of hmc.c:856-874: n_loops, n_measure, m, g, mu, seed."""
    tok = stream.read().split()
    if len(tok) < 6:
        raise ValueError("parameter input needs 6 values: n_loops n_measure m g mu seed")
    return int(tok[0]), int(tok[1]), float(tok[2]), float(tok[3]), float(tok[4]), int(tok[5])


def banner(nt, nx, n_measure, m, g, mu, seed):
    """hmc.c:879-885."""
    return [" ", "++++++++++++++++++++++++++++++++++++++++++",
            " 2D quenched Thirring model, ( %d , %d ) lattice" % (nt, nx),
            " %d updates per measurements" % n_measure, " m %f " % m, " g %f " % g, " mu %f " % mu,
            " Random seed %d" % seed]


def trajectory_lines(obs, chain):
    """obs = the 10 per-chain doubles of tb_hmc_trajectory -> the reference's three lines (hmc.c:701,735,739/743)."""
    p = "[chain %d] " % chain
    return [p + "Start HMC: Sg %s, Smdm %s, Smd %s, Smom %s" % tuple(fmt_g(v) for v in obs[0:4]),
            p + "HMC End, dS %s, Sg %s, Smdm %s, Smd %s, Sm %s" % ((fmt_g(obs[8]),) + tuple(fmt_g(v) for v in obs[4:8])),
            p + ("HMC ACCEPTED" if obs[9] != 0 else "HMC REJECTED")]


def measurement_lines(mag, phase, chain):
    """hmc.c:839-840."""
    p = "[chain %d] " % chain
    return [p + "Magnetisation %s" % fmt_g(mag), p + "Phase %s" % fmt_g(phase)]


def parse_scan(text):
    """'0.2,0.4:0.01,0.1' -> [(g, m), ...] in g-major order (a coupling/mass scan, BASELINE config 5)."""
    gs, ms = text.split(":")
    return [(float(g), float(m)) for g in gs.split(",") for m in ms.split(",")]


def summary_lines(points, cnt, mean, err, names):
    """One line per (g, m) point: the final measurement reduction over all chains of all ranks."""
    out = []
    for p, (g, m) in enumerate(points):
        vals = ", ".join("%s %s +- %s" % (n, fmt_g(mean[p, k]), fmt_g(err[p, k])) for k, n in enumerate(names))
        out.append("[point %d g %s m %s] chains %d, %s" % (p, fmt_g(g), fmt_g(m), int(cnt[p]), vals))
    return out


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--nt", type=int, default=32)
    ap.add_argument("--nx", type=int, default=32)
    ap.add_argument("--chains", type=int, default=1, help="chains per (g, m) point, over all ranks")
    ap.add_argument("--mode", choices=["compat", "adjoint"], default="compat")
    ap.add_argument("--nsteps", type=int, default=10, help="leapfrog steps (hard-coded 10 in hmc.c:708)")
    ap.add_argument("--traj-length", type=float, default=1.0, help="hard-coded 1 in hmc.c:709")
    ap.add_argument("--condensate", type=int, default=0, metavar="NSRC",
                    help="also print the chiral condensate from NSRC stochastic sources (not in the reference's measure())")
    ap.add_argument("--scan", default=None, metavar="G1,G2,..:M1,M2,..",
                    help="coupling/mass scan: every (g, m) pair gets --chains chains; replaces g and m of the parameter input")
    ap.add_argument("--quiet", action="store_true", help="no per-chain lines, only the banner and the final summary")
    ap.add_argument("--backend", default=None, help="torch.distributed backend when launched by torchrun (default nccl)")
    ap.add_argument("--device", type=int, default=None, help="CUDA device (default: LOCAL_RANK modulo the device count)")
    ap.add_argument("--checkpoint", default=None, help="write the gauge fields here at the end (one file per rank)")
    ap.add_argument("--resume", default=None, help="start from this checkpoint instead of the heat bath; the trajectory "
                    "index (it keys the random stream) continues from the one recorded in the file")
    ap.add_argument("--traj-offset", type=int, default=None, help="index of the first trajectory of this run (default: 1, "
                    "or the index recorded in the --resume file); needed to resume a file that does not record it")
    a = ap.parse_args(argv)
    import os

    import thirring2d_b200 as tb
    from thirring2d_b200.shard import chain_range, reduce_observables

    # one process per GPU under torchrun: chains (and with them the scan points) shard block-wise over the ranks,
    # no data-path collective; the only collective is the final measurement reduction
    world, rank = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0"))
    dist = None
    ndev = tb.load_library().tb_device_count()
    device = a.device if a.device is not None else int(os.environ.get("LOCAL_RANK", "0")) % max(ndev, 1)
    if world > 1:
        import torch
        import torch.distributed as dist

        backend = a.backend or "nccl"
        if backend == "nccl":
            torch.cuda.set_device(device)
        dist.init_process_group(backend)

    n_loops, n_measure, m, g, mu, seed = read_parameters(sys.stdin) if world == 1 else _broadcast_parameters(dist, rank)
    points = parse_scan(a.scan) if a.scan else [(g, m)]
    total = a.chains * len(points)
    first, nloc = chain_range(rank, world, total)
    gid = (first + np.arange(nloc)) // a.chains           # scan point of every local chain
    if rank == 0:
        print("\n".join(banner(a.nt, a.nx, n_measure, m, g, mu, seed)))
        if a.scan:
            print(" scan of %d (g, m) points x %d chains over %d rank(s)" % (len(points), a.chains, world))
    mode = tb.MODE_ADJOINT if a.mode == "adjoint" else tb.MODE_REF_COMPAT
    names = ["acceptance", "Magnetisation", "Phase"] + (["Condensate"] if a.condensate > 0 else [])
    acc_sum = np.zeros(nloc)
    meas_sum = np.zeros((nloc, len(names)))   # per chain: sum over the measurements of the run
    n_meas = 0
    if nloc > 0:
        with tb.Context(a.nt, a.nx, nloc, mode, device=device, m=1.0, mu=mu) as ctx:
            ctx.set_params(np.array([points[p][1] for p in gid]), mu)
            ctx.hmc_set_coupling(np.array([points[p][0] for p in gid]))
            ctx.hmc_set_chain_offset(first)
            start = 1
            if a.resume:
                recorded = ctx.checkpoint_read(a.resume if world == 1 else "%s.rank%d" % (a.resume, rank))
                if a.traj_offset is None and recorded == 0:
                    raise SystemExit("--resume: %s does not record a trajectory index; pass --traj-offset (restarting "
                                     "at 1 would replay the random numbers of the first leg)" % a.resume)
                start = a.traj_offset if a.traj_offset is not None else recorded
            else:
                ctx.hmc_heatbath(100, seed=seed)  # hmc.c:927-929
                if a.traj_offset is not None:
                    start = a.traj_offset
            for i in range(start, start + n_loops):
                obs, acc, _ = ctx.hmc_trajectory(a.nsteps, a.traj_length, seed=seed, traj_index=i)
                acc_sum += acc
                if not a.quiet:
                    for c in range(nloc):
                        print("\n".join(trajectory_lines(obs[c], first + c)))
                if i % n_measure == 0:
                    mag, ph = ctx.hmc_measure(20, seed=seed, meas_index=i)
                    n_meas += 1
                    meas_sum[:, 1] += mag
                    meas_sum[:, 2] += ph
                    if not a.quiet:
                        for c in range(nloc):
                            print("\n".join(measurement_lines(mag[c], ph[c], first + c)))
                    if a.condensate > 0:
                        cond, _ = ctx.hmc_condensate(a.condensate, seed=seed, meas_index=i)
                        meas_sum[:, 3] += cond
                        if not a.quiet:
                            for c in range(nloc):
                                print("[chain %d] Condensate %s" % (first + c, fmt_g(cond[c])))
            if a.checkpoint:
                ctx.checkpoint_write(a.checkpoint if world == 1 else "%s.rank%d" % (a.checkpoint, rank),
                                     next_trajectory=start + n_loops)
    # per chain: acceptance over the run and the measurements averaged over the run's trajectories (the summary's
    # error bar is the spread of these chain averages over the chains of a point)
    chain_avg = meas_sum / max(n_meas, 1)
    chain_avg[:, 0] = acc_sum / max(n_loops, 1)
    sys.stdout.flush()
    cnt, mean, err = reduce_observables(chain_avg, gid, len(points), dist=dist,
                                        device=("cuda:%d" % device) if (dist is not None and dist.get_backend() == "nccl") else None)
    if rank == 0:
        print("\n".join(summary_lines(points, cnt, mean, err, names)))
    if dist is not None:
        dist.destroy_process_group()
    return 0


def _broadcast_parameters(dist, rank):
    """Rank 0 reads the parameter input; everybody gets the six values."""
    vals = [read_parameters(sys.stdin)] if rank == 0 else [None]
    dist.broadcast_object_list(vals, src=0)
    return vals[0]


if __name__ == "__main__":
    sys.exit(main())
