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


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--nt", type=int, default=32)
    ap.add_argument("--nx", type=int, default=32)
    ap.add_argument("--chains", type=int, default=1)
    ap.add_argument("--mode", choices=["compat", "adjoint"], default="compat")
    ap.add_argument("--nsteps", type=int, default=10, help="leapfrog steps (hard-coded 10 in hmc.c:708)")
    ap.add_argument("--traj-length", type=float, default=1.0, help="hard-coded 1 in hmc.c:709")
    ap.add_argument("--condensate", type=int, default=0, metavar="NSRC",
                    help="also print the chiral condensate from NSRC stochastic sources (not in the reference's measure())")
    ap.add_argument("--device", type=int, default=0)
    ap.add_argument("--checkpoint", default=None, help="write the gauge fields here at the end")
    ap.add_argument("--resume", default=None, help="start from this checkpoint instead of the heat bath")
    a = ap.parse_args(argv)
    import thirring2d_b200 as tb

    n_loops, n_measure, m, g, mu, seed = read_parameters(sys.stdin)
    print("\n".join(banner(a.nt, a.nx, n_measure, m, g, mu, seed)))
    mode = tb.MODE_ADJOINT if a.mode == "adjoint" else tb.MODE_REF_COMPAT
    with tb.Context(a.nt, a.nx, a.chains, mode, device=a.device, m=m, mu=mu) as ctx:
        ctx.hmc_set_coupling(g)
        if a.resume:
            ctx.checkpoint_read(a.resume)
        else:
            ctx.hmc_heatbath(100, seed=seed)  # hmc.c:927-929
        for i in range(1, n_loops + 1):
            obs, acc, _ = ctx.hmc_trajectory(a.nsteps, a.traj_length, seed=seed, traj_index=i)
            for c in range(a.chains):
                print("\n".join(trajectory_lines(obs[c], c)))
            if i % n_measure == 0:
                mag, ph = ctx.hmc_measure(20, seed=seed, meas_index=i)
                for c in range(a.chains):
                    print("\n".join(measurement_lines(mag[c], ph[c], c)))
                if a.condensate > 0:
                    cond, _ = ctx.hmc_condensate(a.condensate, seed=seed, meas_index=i)
                    for c in range(a.chains):
                        print("[chain %d] Condensate %s" % (c, fmt_g(cond[c])))
        if a.checkpoint:
            ctx.checkpoint_write(a.checkpoint)
    return 0


if __name__ == "__main__":
    sys.exit(main())
