#!/usr/bin/env python
"""bench.py — Dirac applies/s of the batched HMC fermion solve on B200 (one JSON line on rank 0).

    python bench.py --gpus N --steps K --warmup W            (N > 1: launched by torch.distributed.run)
    python bench.py --impl reference ...                     (the reference's CPU implementation, host cores)

A "step" is one pass of the hot path over one batch: fmdm_invert_cg (hmc.c:341-404) for every chain of the
workload from x0 = 0 to ||r||^2 < 1e-30, i.e. 2 Dirac applies (M, M^dagger) + the fused BLAS-1 per CG iteration
per chain.  Workload = BASELINE.json configs[1]: 64x64 lattice, 256 independent chains per GPU, ADJOINT mode,
m = 0.1, mu = 0, quenched-equilibrium links at g = 0.3.  Chains shard over ranks with no
data-path collective (weak scaling: 256 chains per GPU).

  value      whole-job applies/s, inputs resident in HBM, device time of the K solves by CUDA events (the L2 flush
             between steps excluded), max over ranks
  e2e        the same metric through the reference-facing host-buffer C-ABI call (tb_cg_gauge = tb_set_gauge + tb_cg) with
             pinned HOST buffers; H2D of links and sources and D2H of the solutions inside the timed region
  roofline   CG iteration (the 4 fused streaming kernels, or the single resident kernel): algorithmic
             288 B/site/iteration over the device time of the timed solves
  cpu_baseline  the reference's own fmdm_invert_cg (oracle/_ref) on the host cores, bounded sample
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

NT = NX = 64
CHAINS_PER_GPU = 256
MASS, MU, G = 0.1, 0.0, 0.3
BYTES_PER_SITE_ITER = 288  # K1 64 + K2 80 + K3 96 + K4 48
BYTES_PER_SITE_APPLY = 64
METRIC, UNIT = "dirac_applies_per_sec", "applies/s"
HMC_NSTEPS = 40  # m = 0.1 needs 40 leapfrog steps (the reference hard-codes 10, hmc.c:708)


def workload_name(chains):
    return f"cg_solve_{NT}x{NX}_lattice_{chains}_chains_per_gpu_adjoint_m{MASS}_g{G}"


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def synthetic_batch(first_chain, nchains):
    """Inputs of chain c depend only on c (so every rank count gives the same global ensemble).  Same
    generator as oracle/cpu_baseline.py:synthetic_chain, restated here because the product side of the
    bench must not import the oracle."""
    A = np.empty((nchains, NT, NX, 2))
    xi = np.empty((nchains, NT, NX), dtype=np.complex128)
    for i in range(nchains):
        rng = np.random.default_rng(1_000_003 * (first_chain + i + 1))
        A[i] = rng.vonmises(0.0, 2.0 / G, size=(NT, NX, 2))
        xi[i] = rng.normal(size=(NT, NX)) + 1j * rng.normal(size=(NT, NX))
    return A, xi


# ------------------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons through NVML while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.stop_flag, self.samples, self.reasons, self.max_mhz = index, False, [], set(), None
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
        }
        while not self.stop_flag:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.02)

    def result(self):
        self.stop_flag = True
        if self.nv is None or not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["nvml_unavailable"]}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": float(self.max_mhz),
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ------------------------------------------------------------------------------------------------------
def run_cpu_reference(chains_total, procs, flavour="adjoint"):
    """Time the reference's fmdm_invert_cg on `procs` host processes (it is single-threaded: one chain per
    process at a time).  Returns (applies/s, seconds wall, applies, kind)."""
    per = max(1, chains_total // procs)
    cmds = []
    for p in range(procs):
        cmds.append([sys.executable, "-m", "oracle.cpu_baseline", "--nt", str(NT), "--nx", str(NX), "--chains",
                     str(per), "--first", str(p * per), "--m", str(MASS), "--mu", str(MU), "--g", str(G),
                     "--flavour", flavour])
    t0 = time.perf_counter()
    ps = [subprocess.Popen(c, cwd=ROOT, stdout=subprocess.PIPE, text=True) for c in cmds]
    outs = [json.loads(p.communicate()[0].strip().splitlines()[-1]) for p in ps]
    wall = time.perf_counter() - t0
    applies = sum(o["applies"] for o in outs)
    busy = max(o["seconds"] for o in outs)  # solver time of the slowest worker (excludes python start-up)
    return applies / busy, busy, applies, outs[0]["kind"], per * procs, wall


def run_cpu_reference_hmc(procs):
    """One full trajectory of the reference's own driver per host core (oracle/_ref/ref_hmc + the 40-step
    ADJOINT build of hmc.c); returns trajectories/s over all cores, or None when the build is absent."""
    exe = os.path.join(ROOT, "oracle", "_ref", "ref_hmc")
    so = os.path.join(ROOT, "oracle", "_ref", f"libhmcref_{NT}x{NX}_adjoint_ns{HMC_NSTEPS}.so")
    if not (os.path.exists(exe) and os.path.exists(so)):
        return None
    t0 = time.perf_counter()
    ps = [subprocess.Popen([exe, so], stdin=subprocess.PIPE, stdout=subprocess.DEVNULL, text=True) for _ in range(procs)]
    for i, p in enumerate(ps):
        p.stdin.write(f"1\n100\n{MASS}\n{G}\n{MU}\n{4354365264 + 2 * i}\n")
        p.stdin.close()
    for p in ps:
        p.wait()
    return procs / (time.perf_counter() - t0)


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    per_step = cores * 2  # bounded sample: 2 chain solves per core per step (~0.5 s)
    for _ in range(args.warmup):
        run_cpu_reference(cores, cores)
    t_busy, n_applies = 0.0, 0
    for _ in range(args.steps):
        v, busy, applies, kind, nchains, wall = run_cpu_reference(per_step, cores)
        t_busy += busy
        n_applies += applies
    value = n_applies / t_busy
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_busy / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(CHAINS_PER_GPU), "sample": f"{per_step} of the chains per step"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind,
                         "sample": f"{per_step} chain solves per step ({NT}x{NX}, same inputs as the GPU arm), "
                                   f"{cores} single-threaded processes of the reference's fmdm_invert_cg"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------------
def gpu_arm(args):
    import torch
    import torch.distributed as dist

    import thirring2d_b200 as tb

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the product has no CPU path; use --impl reference)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    chains = args.chains
    first = rank * chains  # chain-parallel sharding: rank r owns chains [r*chains, (r+1)*chains)

    stream = torch.cuda.current_stream()
    ctx = tb.Context(NT, NX, chains, tb.MODE_ADJOINT, device=local, m=MASS, mu=MU, stream=stream.cuda_stream)
    if args.solver or args.rows or args.chunk:
        ctx.set_tuning(args.rows, args.chunk, args.solver)

    # synthetic inputs: host (pinned) and device copies
    A_np, xi_np = synthetic_batch(first, chains)
    A_host = torch.from_numpy(A_np).pin_memory()
    n = ctx.vec_doubles
    A_dev = A_host.to(dev)
    xi_canon = torch.from_numpy(xi_np.view(np.float64)).to(dev)
    xi = torch.empty(n, dtype=torch.float64, device=dev)
    b = torch.empty_like(xi)
    x = torch.empty_like(xi)
    ctx.set_gauge_dev(A_dev.data_ptr())
    ctx.pack_dev(xi_canon.data_ptr(), xi.data_ptr())
    ctx.apply_dev(tb.OP_MCONJ, xi.data_ptr(), b.data_ptr())  # b = M~ xi, hmc.c:432
    # host copies of the sources for the e2e leg (canonical layout)
    b_canon = torch.empty_like(xi_canon)
    ctx.unpack_dev(b.data_ptr(), b_canon.data_ptr())
    b_host = b_canon.cpu().pin_memory()
    x_host = torch.empty_like(b_host).pin_memory()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step():
        flush.zero_()  # evict the previous step's state from L2
        ctx.cg_dev(b.data_ptr(), x.data_ptr())
        return ctx.last_solve_ms

    sampler = ClockSampler(local)
    sampler.start()
    for _ in range(max(args.warmup, 3)):
        step()
    info = ctx.cg_result()
    assert np.all(info.status == tb.CG_CONVERGED), info
    iters = info.iters.astype(np.int64)
    applies_per_step = int(2 * iters.sum())

    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ctx.reset_launch_count()
    solve_ms = 0.0
    e0.record()
    for _ in range(args.steps):
        solve_ms += step()
    e1.record()
    barrier()
    launches = ctx.launch_count
    ms_total = e0.elapsed_time(e1)
    # The timed steps re-solve the same sources, so the planned launch (tb_onchip.cuh: TbPlan) has exact iteration counts
    # from the warm-up: its best case.  A first solve, or a solve after the field moved, has no or stale estimates; the
    # plain launch (one CTA per chain, two waves) is what it costs then.
    os.environ["TB_NO_PLAN"] = "1"
    plain_ms = []
    for _ in range(4):
        plain_ms.append(step())
    del os.environ["TB_NO_PLAN"]
    plain_ms = float(np.mean(plain_ms[1:]))
    step()   # planned again, so that x below is the timed solve's
    fp64_peak = ctx.measure_fp64_peak(3)
    fp64_distinct = ctx.measure_fp64_rate(1, 2)   # the same FMAs with every source operand in its own register

    # e2e: the host-buffer C-ABI entry points a reference-side caller binds
    def e2e_step():   # new angles + solve, as at every leapfrog step of momentum_step (hmc.c:504-516)
        ctx.cg_gauge_host_ptr(A_host.data_ptr(), b_host.data_ptr(), x_host.data_ptr())

    for _ in range(2):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_step()
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    x_dev_canon = torch.empty_like(xi_canon)
    ctx.unpack_dev(x.data_ptr(), x_dev_canon.data_ptr())
    assert torch.equal(x_dev_canon.cpu(), x_host), "e2e and device-resident solutions differ"
    # what bounds e2e: the pinned-memory copy rates of this box, measured with the step's own buffers
    pcie = {}
    for name, dst, src in (("h2d_gbs", b_canon, b_host), ("d2h_gbs", x_host, b_canon)):
        c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        dst.copy_(src, non_blocking=True)
        torch.cuda.synchronize()
        c0.record()
        for _ in range(5):
            dst.copy_(src, non_blocking=True)
        c1.record()
        torch.cuda.synchronize()
        pcie[name] = 5 * src.numel() * 8 / (c0.elapsed_time(c1) * 1e-3) / 1e9
    clocks = sampler.result()  # sampled every 20 ms over warm-up, the timed region and the e2e region

    # HMC trajectories/s on the same workload (device-resident update_gauge, hmc.c:671-746, for all chains)
    hmc = None
    if not args.no_hmc:
        ctx.hmc_set_coupling(G)
        ctx.hmc_heatbath(100, seed=1000 + rank)           # main()'s quenched start, hmc.c:927-929
        ctx.hmc_trajectory(HMC_NSTEPS, 1.0, seed=77 + rank, traj_index=0)
        barrier()
        t0 = time.perf_counter()
        acc_sum, it_sum = 0.0, 0
        for k in range(args.traj):
            obs, acc, its = ctx.hmc_trajectory(HMC_NSTEPS, 1.0, seed=77 + rank, traj_index=1 + k)
            acc_sum += float(acc.mean())
            it_sum += its
        torch.cuda.synchronize()
        hmc_s = time.perf_counter() - t0
        hmc = (hmc_s, acc_sum / args.traj, it_sum / (args.traj * chains * (HMC_NSTEPS + 1)))

    t = torch.tensor([ms_total, e2e_s * 1e3, solve_ms, hmc[0] if hmc else 0.0, plain_ms], dtype=torch.float64, device=dev)
    cnt = torch.tensor([applies_per_step, launches], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
    ms_total, e2e_ms, solve_ms, hmc_s, plain_ms = t.tolist()
    applies_all, launches_all = cnt.tolist()

    stream_roof = None
    if rank == 0 and world == 1 and not args.no_extra:
        stream_roof = streaming_roofline(tb, torch, dev, stream)
    other = None
    if not args.no_extra:
        other = other_configs(tb, torch, dist if world > 1 else None, dev, stream, rank, world,
                              cpu=not args.no_cpu_baseline)
    if rank == 0:
        peak, peak_src = measured_peak()
        sites = NT * NX * chains
        max_it = int(iters.max())
        # every chain of a tile streams until the tile's slowest chain converges -> bytes = sum over chains
        alg_bytes_per_step = BYTES_PER_SITE_ITER * NT * NX * float(iters.sum())
        achieved = alg_bytes_per_step * args.steps / (solve_ms * 1e-3) / 1e9
        # device time of the K timed solves (CUDA events around each solve on the context's stream, max over
        # ranks); the L2 flush between steps is not part of the workload (ms_total includes it)
        value = applies_all * args.steps / (solve_ms * 1e-3)
        e2e_value = applies_all * args.steps / (e2e_ms * 1e-3)
        resident = launches <= 2 * args.steps  # one launch per solve => the on-chip resident kernel ran
        launches_per_step = max(launches / args.steps, 1)
        kernel_name = ("resident_wt_kernel (whole batched solve in one launch; r, p in registers, p/Mp exchange in "
                       "shared memory, links and x in tensor memory)"
                       if resident else "CG iteration (dslash, dslash+dot, axpy+norm, xpay)")
        regime = ("cache-resident: state on chip, frac > 1 is expected; HBM traffic = `traffic`"
                  if resident else "streaming")
        traffic, traffic_src = None, None
        tp = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tp):
            tj = json.load(open(tp))
            traffic = tj.get("resident_cg_kernel" if resident else "streaming_iteration")
            traffic_src = (tj.get(("resident_cg_kernel" if resident else "streaming_iteration") + "_source", "")
                           + " -- from a committed ncu --set full capture, NOT measured in this run")
        fp64_tflops = 88 * NT * NX * float(iters.sum()) * args.steps / (solve_ms * 1e-3) / 1e12
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": solve_ms / args.steps,
            "ms_per_step_incl_l2_flush": ms_total / args.steps,
            "ms_per_step_plain_launch": plain_ms,
            "planner_note": "value / ms_per_step: planned launch with the iteration counts of the previous solve of the "
                            "same sources (its best case); ms_per_step_plain_launch: one CTA per chain without a plan, "
                            "what a first solve or a solve on a moved field costs; the hmc figure below has the real mix",
            "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(chains), "lattice": [NT, NX], "chains_per_gpu": chains,
                       "mode": "ADJOINT", "m": MASS, "mu": MU, "g": G, "cg_accuracy": 1e-30,
                       "cg_iters_mean": float(iters.mean()), "cg_iters_max": max_it,
                       "l2": "256 MiB flush buffer written before every timed step, outside the per-step CUDA-event pair"},
            "site_applies_per_sec": value * NT * NX,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src,
                         "peak_source": peak_src, "kernel": kernel_name, "regime": regime,
                         # the solver kernel runs once per step (the one-block plan_kernel in front of it moves no
                         # lattice data); streaming: per launch of the 3 kernels of an iteration
                         "algorithmic_bytes_per_launch": alg_bytes_per_step if resident
                         else alg_bytes_per_step / launches_per_step,
                         "algorithmic_bytes_per_site_iteration": BYTES_PER_SITE_ITER,
                         "us_per_iteration": solve_ms * 1e3 / args.steps / max_it, "sites": sites,
                         # what actually bounds the on-chip kernel (ncu: FP64 pipe), for orientation only: 34
                         # useful flops per site and apply + 20 for the fused BLAS-1
                         "fp64": {"achieved_tflops": fp64_tflops, "measured_peak_tflops": fp64_peak,
                                  "frac_of_measured": fp64_tflops / fp64_peak if fp64_peak > 0 else None,
                                  "measured_rate_distinct_operands_tflops": fp64_distinct,
                                  "flops_per_site_iteration": 88,
                                  "peak_source": "tb_measure_fp64_peak: independent DFMA chains on every SM, best of 3 "
                                                 "launches, in this run"}},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(A_host.nbytes + b_host.nbytes),
                    "d2h_bytes_per_step": int(x_host.nbytes), "ms_per_step": e2e_ms / args.steps,
                    "pinned_copy_rates_measured": pcie,
                    "note": "per sub-batch of chains: H2D of angles and sources -> ONE kernel that reads and writes the host's "
                            "layout and builds its links -> D2H; chains are not split on this path, so 256 chains on 148 SMs "
                            "cost two waves of a 0.9 ms solve behind the arrival of the 108th chain (floor 2.05 ms per step); at "
                            "N = 8 the ranks share one host memory system (pinned_copy_rates_measured falls from 55 to 20 GB/s)"},
            "gpu_launches": int(launches_all),
            "clocks": clocks,
        }
        if hmc is not None:
            line["hmc"] = {"traj_per_sec": world * chains * args.traj / hmc_s, "unit": "trajectories/s",
                           "ms_per_batched_trajectory": 1e3 * hmc_s / args.traj, "nsteps": HMC_NSTEPS,
                           "traj_length": 1.0, "acceptance": hmc[1], "cg_iters_per_solve": hmc[2],
                           "note": "update_gauge as coded (hmc.c:671-746) for every chain, device-resident; "
                                   "nsteps = 40 because the hard-coded 10 has zero acceptance at m = 0.1 "
                                   "(SURVEY Appendix C)"}
        if stream_roof is not None:
            line["roofline_streaming"] = stream_roof
        if other is not None:
            line["other_configs"] = other
        if world == 1 and not args.no_cpu_baseline:
            cores = os.cpu_count() or 1
            v, busy, applies, kind, nch, wall = run_cpu_reference(cores * 8, cores)
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": kind,
                                    "sample": f"{nch} chain solves of the same workload ({NT}x{NX}, m={MASS}), "
                                              f"{cores} single-threaded processes, {busy:.1f} s"}
            line["parity_checked"] = parity_check(torch, tb, ctx, x, A_np, b_host, info, chains)
            # also one core, and the flags the reference ships with (Makefile:2, no optimisation)
            v1 = run_cpu_reference(4, 1)
            line["cpu_baseline"]["one_core"] = {"value": v1[0], "sample": f"{v1[4]} chain solves, -O3"}
            if os.path.exists(os.path.join(ROOT, "oracle", "_ref", f"libhmcref_{NT}x{NX}_adjoint_shipped.so")):
                vs = run_cpu_reference(2, 1, "adjoint_shipped")
                line["cpu_baseline"]["one_core_shipped_flags"] = {
                    "value": vs[0], "sample": f"{vs[4]} chain solves, the reference's own CFLAGS (-std=c99 -g, no -O)"}
            if hmc is not None:
                tps = run_cpu_reference_hmc(cores)
                if tps is not None:
                    line["hmc"]["cpu_baseline"] = {"traj_per_sec": tps, "cores": cores, "kind": "reference",
                                                   "sample": f"1 trajectory per core of the reference's own hmc.c "
                                                             f"driver ({NT}x{NX}, m={MASS}, {HMC_NSTEPS} steps)"}
        print(json.dumps(line), flush=True)
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


def parity_check(torch, tb, ctx, x, A_np, b_host, info, chains):
    """Outside the timed region: three of the headline chains against the CPU checker (the oracle port, bit-identical
    to the compiled reference) on the very inputs and solutions of the timed solves."""
    from oracle.pyoracle import MODE_ADJOINT, Oracle   # the checker, never the thing measured

    orc = Oracle()
    xc = torch.empty_like(x)
    ctx.unpack_dev(x.data_ptr(), xc.data_ptr())
    xh = xc.cpu().numpy().view(np.complex128).reshape(chains, NT, NX)
    bh = b_host.numpy().view(np.complex128).reshape(chains, NT, NX)
    out = {"chains": [], "iters_gpu": [], "iters_reference": [], "solution_rel_l2": []}
    for c in sorted({0, chains // 2, chains - 1}):
        xo, st, it, rr = orc.fmdm_invert_cg(bh[c], A_np[c], MASS, MU, MODE_ADJOINT)
        out["chains"].append(int(c))
        out["iters_gpu"].append(int(info.iters[c]))
        out["iters_reference"].append(int(it))
        out["solution_rel_l2"].append(float(np.linalg.norm(xh[c] - xo) / np.linalg.norm(xo)))
    out["ok"] = bool(all(abs(a - b) <= 1 for a, b in zip(out["iters_gpu"], out["iters_reference"]))
                     and max(out["solution_rel_l2"]) <= 1e-12)
    out["tolerance"] = "iterations +-1, solution 1e-12 l2-relative (north_star, SURVEY Appendix C)"
    return out


def streaming_roofline(tb, torch, dev, stream):
    """HBM-bound evidence on a working set larger than L2: the streaming CG (3 fused kernels per iteration; the two stencil
    passes TMA-staged, 16-chain tiles through tensor maps) on 256x256 x 64 chains (470 MB of CG state), fixed 150
    iterations, device time by CUDA events."""
    nt = nx = 256
    chains, iters = 64, 150
    peak, _ = measured_peak()
    ctx = tb.Context(nt, nx, chains, tb.MODE_ADJOINT, device=dev.index or 0, m=0.01, mu=0.0, stream=stream.cuda_stream)
    ctx.set_tuning(0, 0, 1)
    ctx.set_cg(1e-30, iters + 1)
    g = torch.Generator(device=dev).manual_seed(7)
    A = (torch.rand(chains * nt * nx * 2, dtype=torch.float64, device=dev, generator=g) - 0.5) * (2 * np.pi)
    ctx.set_gauge_dev(A.data_ptr())
    n = ctx.vec_doubles
    b = torch.randn(n, dtype=torch.float64, device=dev, generator=g)
    x = torch.empty_like(b)
    ctx.cg_dev(b.data_ptr(), x.data_ptr())
    ms = []
    for _ in range(3):
        ctx.cg_dev(b.data_ptr(), x.data_ptr())
        ms.append(ctx.last_solve_ms)
    info = ctx.cg_result()
    it = int(info.iters.max())
    kernels = ctx.streaming_info()
    ctx.close()
    t = min(ms) * 1e-3
    # the fused ADJOINT iteration moves 240 B per site (K1 64 + the fused M^dagger / update pass 128 + K4 48): that is
    # `achieved`; SURVEY 8(d)'s accounting of the unfused design (288 B) is reported beside it
    ach = 240 * nt * nx * chains * it / t / 1e9
    ach288 = BYTES_PER_SITE_ITER * nt * nx * chains * it / t / 1e9
    return {"bound": "hbm", "kernel": "streaming CG iteration (dslash+|Mp|^2, dslash^dagger+axpy+norm, xpay)",
            "stencil_kernels": {0: "register-marching", 1: "TMA-staged, whole-batch tiles", 2: "TMA-staged, 16-chain tiles "
                                "through tensor maps"}[kernels[0]] + f", tile {kernels[1]} chains x {kernels[2]} sites, "
                               f"{kernels[3]} rows per block",
            "workload": f"{nt}x{nx} x {chains} chains, {it} iterations, working set 470 MB > L2",
            "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "bytes_per_site_iteration": 240,
            "achieved_288B_definition": ach288, "frac_288B_definition": ach288 / peak,
            "us_per_iteration": t * 1e6 / it}


def slab_config(tb, torch, dist, dev, stream, rank, world, n, peak, iters=200, m=0.05):
    """One n x n lattice, `iters` CG iterations of fmdm_invert_cg, on `world` GPUs as t-slabs (world == 1: the
    streaming solver).  Device time of the solve by CUDA events, max over ranks, best of 3 after one warm-up."""
    local = dev.index or 0
    gen = torch.Generator(device=dev).manual_seed(5)   # the same stream on every rank: every rank draws the global field
    A = ((torch.rand(n * n * 2, dtype=torch.float64, device=dev, generator=gen) - 0.5) * (2 * np.pi)).view(n, n, 2)
    b = torch.randn(n * n * 2, dtype=torch.float64, device=dev, generator=gen).view(n, n, 2)

    def single():
        ctx = tb.Context(n, n, 1, tb.MODE_ADJOINT, device=local, m=m, mu=0.0, stream=stream.cuda_stream)
        ctx.set_cg(1e-30, iters + 1)
        ctx.set_tuning(0, 0, 1)
        ctx.set_gauge_dev(A.data_ptr())
        x = torch.empty_like(b)
        ms = []
        for _ in range(4):
            ctx.cg_dev(b.data_ptr(), x.data_ptr())
            ms.append(ctx.last_solve_ms)
        it = int(ctx.cg_result().iters.max())
        ctx.close()
        return min(ms[1:]), it, float((x * x).sum())

    if world == 1:
        ms, it, _ = single()
        return {"us_per_cg_iteration": ms * 1e3 / it, "dirac_applies_per_sec": 2 * it / (ms * 1e-3),
                "site_applies_per_sec": 2 * it * n * n / (ms * 1e-3), "iterations": it, "gpus": 1,
                "hbm_frac_288B_definition": BYTES_PER_SITE_ITER * n * n * it / (ms * 1e-3) / 1e9 / peak,
                "hbm_frac_real_traffic_240B": 240 * n * n * it / (ms * 1e-3) / 1e9 / peak}

    ntl = n // world
    A_l = A[rank * ntl:(rank + 1) * ntl].contiguous()
    b_l = b[rank * ntl:(rank + 1) * ntl].contiguous()
    x_l = torch.empty_like(b_l)
    ctx = tb.Context(n, n, 1, tb.MODE_ADJOINT, device=local, m=m, mu=0.0, stream=stream.cuda_stream,
                     slab_rank=rank, slab_nranks=world)
    ctx.slab_setup(dist)
    ctx.set_cg(1e-30, iters + 1)
    ctx.set_gauge_dev(A_l.data_ptr())
    ctx.synchronize()
    dist.barrier()

    def timed(env=None):
        if env:
            os.environ[env] = "1"
        ms = []
        for _ in range(4):
            dist.barrier()
            ctx.cg_dev(b_l.data_ptr(), x_l.data_ptr())
            ms.append(ctx.last_solve_ms)
        if env:
            del os.environ[env]
        t = torch.tensor([min(ms[1:])], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    ms_multi = timed("TB_NO_PERSIST")   # one kernel per pass, flags between kernels
    ms = timed()                        # default: the whole solve in one launch per GPU where the slab allows it
    it = int(ctx.cg_result().iters.max())
    chk = torch.tensor([float((x_l * x_l).sum())], dtype=torch.float64, device=dev)
    dist.all_reduce(chk, op=dist.ReduceOp.SUM)
    ctx.close()
    one = torch.zeros(3, dtype=torch.float64, device=dev)
    if rank == 0:   # the same problem on one GPU, in this run
        ms1, it1, chk1 = single()
        one = torch.tensor([ms1, it1, chk1], dtype=torch.float64, device=dev)
    dist.broadcast(one, 0)
    ms1, it1, chk1 = one.tolist()
    best = min(ms, ms_multi)
    return {"us_per_cg_iteration": best * 1e3 / it, "gpus": world, "iterations": it,
            "us_per_cg_iteration_one_launch_solve": ms * 1e3 / it,
            "us_per_cg_iteration_multi_kernel": ms_multi * 1e3 / it,
            "us_per_cg_iteration_1_gpu": ms1 * 1e3 / it1, "speedup_vs_1_gpu": (ms1 / it1) / (best / it),
            "dirac_applies_per_sec": 2 * it / (best * 1e-3), "site_applies_per_sec": 2 * it * n * n / (best * 1e-3),
            "x_norm2_rel_diff_vs_1_gpu": abs(chk.item() - chk1) / chk1,
            "rows_per_gpu": ntl, "hbm_frac_real_traffic_240B_per_gpu": 240 * n * n / world * it / (best * 1e-3) / 1e9 / peak,
            "note": "strong scaling of ONE lattice: t-slabs, peer-memory halos and all-reduces (no NCCL on the data "
                    "path); device time, max over ranks"}


def cpu_apply_rate(nt, nx, m, g, iters):
    """Dirac applies/s of the reference's CPU path on ONE host core at a lattice size where a full solve would take
    minutes to hours (SURVEY 8(c) caveat 4): `iters` CG iterations of the reference's own fmdm_invert_cg, compiled from
    the unmodified hmc.c with CG_MAX_ITER rewritten to iters + 1 (oracle/build_ref.sh, libhmcref_*_it<N>.so; max-iter is
    silent, hmc.c:364); the oracle port (bit-identical, tests/test_oracle_pinned.py) when that object is absent."""
    p = subprocess.run([sys.executable, "-m", "oracle.cpu_baseline", "--nt", str(nt), "--nx", str(nx), "--chains", "1",
                        "--m", str(m), "--mu", "0.0", "--g", str(g), "--flavour", f"adjoint_it{iters + 1}",
                        "--max-iter", str(iters + 1)], cwd=ROOT, capture_output=True, text=True)
    try:
        o = json.loads(p.stdout.strip().splitlines()[-1])
    except Exception:
        return {"error": (p.stderr or p.stdout)[-300:]}
    return {"dirac_applies_per_sec": o["applies"] / o["seconds"], "cores": 1, "kind": o["kind"],
            "sample": f"{o['iters']} CG iterations of one {nt}x{nx} chain, {o['seconds']:.1f} s"}


def other_configs(tb, torch, dist, dev, stream, rank, world, cpu=True):
    """The remaining BASELINE.json configurations, one bounded measurement each (not the headline; parity for
    these shapes is in tests/).  Chains / parameter points shard over ranks like the headline workload: every
    figure is whole-job units over the slowest rank's time."""
    out = {}
    local = dev.index or 0

    def max_over_ranks(ms):
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        if dist is not None:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def total_over_ranks(v):
        t = torch.tensor([float(v)], dtype=torch.float64, device=dev)
        if dist is not None:
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    def solve_once(ctx, n, solver):
        g = torch.Generator(device=dev).manual_seed(11 + rank)
        xi = torch.randn(n, dtype=torch.float64, device=dev, generator=g)
        b, x = torch.empty_like(xi), torch.empty_like(xi)
        ctx.apply_dev(tb.OP_MCONJ, xi.data_ptr(), b.data_ptr())
        ctx.set_tuning(0, 0, solver)
        ctx.cg_dev(b.data_ptr(), x.data_ptr())  # warm-up (graph build for the streaming solver)
        ctx.cg_dev(b.data_ptr(), x.data_ptr())
        info = ctx.cg_result()
        return ctx.last_solve_ms, info

    # configs[2]: 256x256, light mass (ill-conditioned CG), chains sharded over the GPUs.  SURVEY 8(d): g = 1,
    # m = 0.01 (about 3 100 iterations), 8 chains per GPU.
    nt = nx = 256
    chains = 8
    ctx = tb.Context(nt, nx, chains, tb.MODE_ADJOINT, device=local, m=0.01, mu=0.0, stream=stream.cuda_stream)
    ctx.hmc_set_coupling(1.0)
    ctx.hmc_heatbath(100, seed=300 + rank)
    kind, in_flight = ctx.solver_info()
    ms, info = solve_once(ctx, ctx.vec_doubles, 0)
    ms_stream, _ = solve_once(ctx, ctx.vec_doubles, 1)
    # trajectories/s of this configuration: update_gauge as coded (10 leapfrog steps, hmc.c:708), device-resident
    ctx.set_tuning(0, 0, 0)
    ctx.hmc_trajectory(10, 1.0, seed=310 + rank, traj_index=0)   # warm-up
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    _, _, traj_its = ctx.hmc_trajectory(10, 1.0, seed=310 + rank, traj_index=1)
    torch.cuda.synchronize()
    traj_s = max_over_ranks((time.perf_counter() - t0) * 1e3) * 1e-3
    ctx.close()
    ms = max_over_ranks(ms)
    applies = total_over_ranks(2 * int(info.iters.astype(np.int64).sum()))
    out["256x256_m0.01_g1_8_chains_per_gpu"] = {
        "dirac_applies_per_sec": applies / (ms * 1e-3), "site_applies_per_sec": applies * nt * nx / (ms * 1e-3),
        "ms_per_batched_solve": ms, "cg_iters_mean": float(info.iters.mean()),
        "converged": bool(np.all(info.status == tb.CG_CONVERGED)),
        "solver": {0: "streaming", 1: "on-chip (CTA per chain)", 2: "on-chip (16-CTA cluster per chain)"}[kind],
        "chains_in_flight": in_flight, "ms_streaming_solver": max_over_ranks(ms_stream),
        "hmc": {"traj_per_sec": world * chains / traj_s, "ms_per_batched_trajectory": 1e3 * traj_s, "nsteps": 10,
                "cg_iters_per_solve": float(traj_its) / (chains * 11)}}
    if world == 1 and cpu:
        out["256x256_m0.01_g1_8_chains_per_gpu"]["cpu_baseline"] = cpu_apply_rate(256, 256, 0.01, 1.0, 40)

    # configs[4]: coupling/mass scan, 32 (g, m) points x 64 chains on 128x128, chiral condensate measurement with
    # N_src = 20 stochastic sources per configuration (hmc.c:798) through fm_invert_cg (hmc.c:408-414).  The SAME 2 048
    # chains at every N (strong scaling of the fixed scan).  A chain's cost goes like its CG iteration count (~ 1/m):
    # chains are dealt to the ranks by expected cost (shard.deal_by_cost, longest first) and every rank runs its long
    # solves first, so that the tail of a batched solve on the cluster solver is made of the cheap chains.
    from thirring2d_b200.shard import deal_by_cost, reduce_observables

    nt = nx = 128
    per_point = 64
    pts = [(g, m) for g in (0.2, 0.3, 0.4, 0.5, 0.6, 0.7, 0.8, 1.0) for m in (0.01, 0.03, 0.1, 0.3)]
    point_of = np.repeat(np.arange(len(pts)), per_point)            # global chain -> scan point
    cost = np.array([30.0 / pts[p][1] for p in point_of])           # expected iterations (measured: 3 000 at m = 0.01)
    mine = deal_by_cost(cost, world)[rank]
    chains = int(mine.size)
    g_arr = np.array([pts[point_of[c]][0] for c in mine])
    m_arr = np.array([pts[point_of[c]][1] for c in mine])
    ctx = tb.Context(nt, nx, chains, tb.MODE_ADJOINT, device=local, stream=stream.cuda_stream)
    ctx.set_params(m_arr, 0.0)
    ctx.hmc_set_coupling(g_arr)
    ctx.hmc_heatbath(100, seed=500 + rank)
    kind, in_flight = ctx.solver_info()
    nsrc = 20
    ctx.hmc_condensate(nsrc=1, seed=1, meas_index=0)  # warm-up
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    t0 = time.perf_counter()
    cond, its = ctx.hmc_condensate(nsrc=nsrc, seed=1, meas_index=1)
    torch.cuda.synchronize()
    my_sec = time.perf_counter() - t0
    sec = max_over_ranks(my_sec * 1e3) * 1e-3
    busy = total_over_ranks(my_sec) / world
    ctx.close()
    applies = total_over_ranks(2 * its + chains * nsrc)
    cnt, mean, err = reduce_observables(cond[:, None], point_of[mine], len(pts), dist=dist, device=dev)
    out["128x128_scan_32_points_x_64_chains_condensate"] = {
        "condensate_sources_per_sec": len(point_of) * nsrc / sec, "dirac_applies_per_sec": applies / sec,
        "seconds": sec, "mean_over_ranks_seconds": busy, "load_balance": busy / sec, "scaling": "strong",
        "chains_total": int(len(point_of)), "chains_on_rank0": chains, "nsrc": nsrc,
        "ms_per_source_batch": 1e3 * sec / nsrc,
        "points": [{"g": pts[p][0], "m": pts[p][1], "chains": int(cnt[p]), "condensate": float(mean[p, 0]),
                    "stderr": float(err[p, 0])} for p in range(len(pts))],
        "finite": bool(np.all(np.isfinite(mean))),
        "solver": {0: "streaming", 1: "on-chip (CTA per chain)", 2: "on-chip (4-CTA cluster per chain)"}[kind],
        "chains_in_flight": in_flight,
        "dealing": "shard.deal_by_cost: expected cost 1/m, longest-processing-time rule, long solves first on each rank"}

    # SURVEY 8(f) row 3, family B (vec_ops.c): measure_propagator's batch of point sources on one occupation mask
    # (fermionbag.c:389-435) through cg_propagator, host buffers in and out; reference = its own vec_ops.c on one core
    if world == 1:
        from_b = subprocess.run([sys.executable, "-m", "oracle.cpu_baseline", "--nt", "64", "--nx", "64", "--m", "0.1",
                                 "--mu", "0.0", "--family-b", "8"], cwd=ROOT, capture_output=True, text=True)
        nt = nx = 64
        nsrc = 2 * nx
        rng = np.random.default_rng(4242)   # same mask and sources as oracle/cpu_baseline.py:family_b_inputs
        field = (rng.random((nt, nx)) < 0.1).astype(np.int32)
        sites = [(t, x) for t in range(nt) for x in range(nx) if field[t, x] == 0][:nsrc]
        src = np.zeros((nsrc, nt, nx))
        for i, (t, x) in enumerate(sites):
            src[i, t, x] = 1.0
        ctx = tb.Context(nt, nx, nsrc, tb.MODE_ADJOINT, device=local, m=0.1, mu=0.0, stream=stream.cuda_stream)
        ctx.set_occupancy(field)   # one (NT, NX) field shared by the whole multi-RHS batch
        kind = "on-chip real CG (one CTA per source, tb_real.cu)"
        # host buffers as the reference's caller holds them (pageable numpy arrays of real vectors) ...
        ctx.cg_propagator(src)
        t0 = time.perf_counter()
        prop, info = ctx.cg_propagator(src)
        sec = time.perf_counter() - t0
        # ... and pinned ones (cudaHostAlloc), which is what the interposed alloc_vector can hand the driver
        src_pin = torch.from_numpy(src).pin_memory()
        out_pin = torch.empty_like(src_pin).pin_memory()
        ctx.cg_real_host_ptr(src_pin.data_ptr(), out_pin.data_ptr(), propagator=True)
        t0 = time.perf_counter()
        for _ in range(5):
            ctx.cg_real_host_ptr(src_pin.data_ptr(), out_pin.data_ptr(), propagator=True)
        sec_pin = (time.perf_counter() - t0) / 5
        dev_ms = ctx.last_solve_ms
        same = bool(np.array_equal(out_pin.numpy(), prop))
        ctx.close()
        fb = {"propagators_per_sec": nsrc / sec_pin, "propagators_per_sec_pageable_numpy": nsrc / sec, "sources": nsrc,
              "ms_per_batch": 1e3 * sec_pin, "ms_per_batch_pageable_numpy": 1e3 * sec, "ms_on_device_incl_copies": dev_ms,
              "h2d_bytes": int(src.nbytes), "d2h_bytes": int(src.nbytes), "pinned_equals_pageable": same,
              "cg_iters_mean": float(info.iters.mean()), "converged": bool(np.all(info.status == tb.CG_CONVERGED)),
              "solver": kind + ", real 8-byte vectors end to end"}
        try:
            o = json.loads(from_b.stdout.strip().splitlines()[-1])
            fb["cpu_baseline"] = {"propagators_per_sec": o["sources"] / o["seconds"], "cores": 1, "kind": o["kind"],
                                  "sample": f"{o['sources']} point sources through the reference's cg_propagator"}
        except Exception:
            pass
        out["family_b_64x64_point_source_propagators"] = fb

    # configs[0]: the reference's own driver (hmc.c UNMODIFIED) with the shipped parameters, once on the CPU and once
    # on top of the library through symbol interposition: wall clock of the whole process
    if world == 1:
        launcher = os.path.join(ROOT, "thirring2d_b200", "hmc_b200")
        ref_exe = os.path.join(ROOT, "oracle", "_ref", "ref_hmc")
        drop = {}
        for name, so, mode, params in (
                ("32x32_shipped_parameter_file_m100", "libhmcref_32x32_compat.so", "compat",
                 "40\n40\n100\n0.3\n0.1\n4354365264\n"),
                ("32x32_adjoint_m0.1_40_steps", "libhmcref_32x32_adjoint_ns40.so", "adjoint",
                 "6\n100\n0.1\n0.3\n0.0\n4354365264\n")):   # no measure(): its test_conjugate aborts a true adjoint
            sop = os.path.join(ROOT, "oracle", "_ref", so)
            if not (os.path.exists(launcher) and os.path.exists(ref_exe) and os.path.exists(sop)):
                continue
            ntraj = int(params.split("\n")[0])
            zero = "0\n" + params.split("\n", 1)[1]   # same start (seeding, 100 heat-bath sweeps), no trajectory

            nsteps_env = dict(os.environ, THIRRING_NSTEPS="40" if "ns40" in so else "10")

            def wall(cmd, inp):
                """Seconds inside the driver's main (printed by both launchers: CUDA start-up and dlopen excluded)."""
                t0 = time.perf_counter()
                p = subprocess.run(cmd, input=inp, capture_output=True, text=True, env=nsteps_env)
                dt = time.perf_counter() - t0
                for ln in p.stderr.splitlines():
                    if ln.startswith("hmc_main_seconds="):
                        dt = float(ln.split("=")[1])
                return dt, p

            gpu_cmd, cpu_cmd = [launcher, sop, "32", "32", mode], [ref_exe, sop]
            wall(gpu_cmd, zero)   # warm the driver / page in the library
            tg0 = min(wall(gpu_cmd, zero)[0] for _ in range(2))
            tg, pg = wall(gpu_cmd, params)
            tc0 = wall(cpu_cmd, zero)[0]
            tc, pc = wall(cpu_cmd, params)
            # optional coarse override (libthirring_hmc_coarse.so): update_gauge = one device-resident trajectory
            co_cmd = gpu_cmd + ["0", "coarse"]
            tk0 = min(wall(co_cmd, zero)[0] for _ in range(2))
            tk, pk = wall(co_cmd, params)
            drop[name] = {"trajectories": ntraj,
                          "ms_per_trajectory_on_the_library": 1e3 * max(tg - tg0, 0.0) / ntraj,
                          "ms_per_trajectory_reference_cpu": 1e3 * max(tc - tc0, 0.0) / ntraj,
                          "driver_seconds_on_the_library": tg, "driver_seconds_reference_cpu": tc,
                          "ms_per_trajectory_coarse_override": 1e3 * max(tk - tk0, 0.0) / ntraj,
                          "stdout_identical_coarse_override": pk.stdout == pc.stdout,
                          "stdout_identical": pg.stdout == pc.stdout,
                          "served": [ln for ln in pg.stderr.splitlines() if ln.startswith("hmc_b200:")][-1:],
                          "served_coarse_override": [ln for ln in pk.stderr.splitlines() if ln.startswith("hmc_b200:")][-1:],
                          "note": "one chain; seconds inside the driver's main(), per trajectory = (N-trajectory run) - "
                                  "(0-trajectory run): CUDA start-up and the common heat-bath start are excluded"}
        if drop:
            out["unmodified_reference_driver_single_chain"] = drop

    # configs[3]: 2048x2048 (and 4096x4096) single lattice, CG with a fixed 200 iterations (hmc.c:364-395).  N = 1: the
    # streaming solver.  N > 1: the lattice is cut into N t-slabs, one per rank (tb_create_slab): halo rows are peer
    # loads over NVLink, the two dot products per iteration one-shot peer-store all-reduces; the same global field and
    # source on every N, and rank 0 also solves the whole lattice alone so that the speed-up is measured in this run.
    peak, _ = measured_peak()
    for n in (2048, 4096):
        key = f"{n}x{n}_single_lattice" + ("_1_gpu" if world == 1 else "_slab")
        try:
            out[key] = slab_config(tb, torch, dist, dev, stream, rank, world, n, peak)
        except Exception as e:   # a dead peer or a launch error must not cost the headline line
            out[key] = {"error": f"{type(e).__name__}: {e}"[:300]}
            break
    if world == 1 and cpu and "error" not in out["2048x2048_single_lattice_1_gpu"]:
        out["2048x2048_single_lattice_1_gpu"]["cpu_baseline"] = cpu_apply_rate(2048, 2048, 0.05, 1.0, 2)
    if world == 1:
        try:
            out["1024x1024_16_sources_on_one_gauge_field"] = multi_rhs_config(tb, torch, dev, stream, peak)
        except Exception as e:
            out["1024x1024_16_sources_on_one_gauge_field"] = {"error": f"{type(e).__name__}: {e}"[:300]}
    return out


def multi_rhs_config(tb, torch, dev, stream, peak, n=1024, nsrc=16, iters=200):
    """Multi-RHS CG on ONE gauge field (the sources of fermion_phase, hmc.c:794-815; SURVEY 8(d): 32 + 32 / N_src bytes of
    links per site and apply): the field uploaded once per source (tb_set_gauge_dev) against tb_set_gauge_shared_dev, where
    the staged kernels read one link per site.  Fixed number of iterations, device time."""
    res = {}
    g = torch.Generator(device=dev).manual_seed(11)
    A1 = (torch.rand(n * n * 2, dtype=torch.float64, device=dev, generator=g) - 0.5) * (2 * np.pi)
    for shared in (False, True):
        ctx = tb.Context(n, n, nsrc, tb.MODE_ADJOINT, device=dev.index or 0, m=0.01, mu=0.0, stream=stream.cuda_stream)
        ctx.set_tuning(0, 0, 1)
        ctx.set_cg(1e-30, iters + 1)
        if shared:
            ctx.set_gauge_shared_dev(A1.data_ptr())
        else:
            A = A1.view(1, -1).expand(nsrc, -1).contiguous()
            ctx.set_gauge_dev(A.data_ptr())
            del A
        b = torch.randn(ctx.vec_doubles, dtype=torch.float64, device=dev, generator=g)
        x = torch.empty_like(b)
        ctx.cg_dev(b.data_ptr(), x.data_ptr())
        ms = []
        for _ in range(3):
            ctx.cg_dev(b.data_ptr(), x.data_ptr())
            ms.append(ctx.last_solve_ms)
        it = int(ctx.cg_result().iters.max())
        kern = ctx.streaming_info()
        ctx.close()
        del b, x
        res["shared" if shared else "per_source"] = (min(ms) * 1e3 / it, kern)
    sites = n * n * nsrc
    us_s, us_p = res["shared"][0], res["per_source"][0]
    return {"workload": f"{n}x{n}, {nsrc} sources on one gauge field, {iters} CG iterations, streaming solver",
            "us_per_iteration_field_per_source": us_p, "us_per_iteration_shared_field": us_s, "speedup": us_p / us_s,
            "bytes_per_site_source_iteration": {"per_source": 224, "shared": 224 - 64 + 64 / nsrc},
            "frac_of_hbm_peak_shared": (224 - 64 + 64 / nsrc) * sites / (us_s * 1e-6) / 1e9 / peak,
            "rows_per_block": {"per_source": res["per_source"][1][3], "shared": res["shared"][1][3]}}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--chains", type=int, default=CHAINS_PER_GPU)
    ap.add_argument("--rows", type=int, default=0)
    ap.add_argument("--chunk", type=int, default=0)
    ap.add_argument("--solver", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the streaming-path roofline measurement")
    ap.add_argument("--no-hmc", action="store_true", help="skip the trajectories/s measurement")
    ap.add_argument("--traj", type=int, default=2, help="timed trajectories for the hmc figure")
    args = ap.parse_args()
    if args.impl == "reference":
        reference_arm(args)
    else:
        gpu_arm(args)


if __name__ == "__main__":
    main()
