"""Config 4 of BASELINE.json: one NT x NX lattice, CG slab-decomposed over the GPUs of the box.
Launch:  python -m torch.distributed.run --nproc-per-node P --master-addr 127.0.0.1 tools/slab_bench.py [--size 2048]
Prints one JSON line on rank 0: microseconds per CG iteration (fixed iteration count), max over ranks."""
import argparse
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import thirring2d_b200 as tb


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", type=int, default=2048)
    ap.add_argument("--iters", type=int, default=200)
    ap.add_argument("--m", type=float, default=0.05)
    ap.add_argument("--chunk", type=int, default=0)
    ap.add_argument("--rows", type=int, default=0)
    a = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    n, ntl = a.size, a.size // world
    stream = torch.cuda.current_stream()
    if world > 1:
        ctx = tb.Context(n, n, 1, tb.MODE_ADJOINT, device=local, m=a.m, mu=0.0, stream=stream.cuda_stream,
                         slab_rank=rank, slab_nranks=world)
        ctx.slab_setup(dist)
    else:
        ctx = tb.Context(n, n, 1, tb.MODE_ADJOINT, device=local, m=a.m, mu=0.0, stream=stream.cuda_stream)
    ctx.set_tuning(a.rows, a.chunk, 1)
    ctx.set_cg(1e-30, a.iters + 1)
    # every rank draws the global field from the same seed and keeps its slab
    g = torch.Generator(device="cpu").manual_seed(11)
    A = (torch.rand(n, n, 2, dtype=torch.float64, generator=g) - 0.5) * (2 * np.pi)
    b = torch.randn(n, n, 2, dtype=torch.float64, generator=g)
    A_l = A[rank * ntl:(rank + 1) * ntl].contiguous().to(dev)
    b_l = b[rank * ntl:(rank + 1) * ntl].contiguous().to(dev)
    x_l = torch.empty_like(b_l)
    ctx.set_gauge_dev(A_l.data_ptr())
    ctx.synchronize()
    if world > 1:
        dist.barrier()
    ms = []
    for _ in range(4):
        if world > 1:
            dist.barrier()
        ctx.cg_dev(b_l.data_ptr(), x_l.data_ptr())
        ms.append(ctx.last_solve_ms)
    info = ctx.cg_result()
    t = torch.tensor([min(ms[1:])], dtype=torch.float64, device=dev)
    chk = torch.tensor([float((x_l.double() ** 2).sum())], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(chk, op=dist.ReduceOp.SUM)
    if rank == 0:
        it = int(info.iters.max())
        us = t.item() * 1e3 / it
        print(json.dumps({"lattice": [n, n], "gpus": world, "iters": it, "us_per_iteration": round(us, 2),
                          "site_iters_per_s": n * n * it / (t.item() * 1e-3),
                          "algorithmic_GBs_total": round(288.0 * n * n / us / 1e3, 1),
                          "x_norm2": chk.item()}), flush=True)
    ctx.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
