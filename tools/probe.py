"""Quick device-side timing probe (development aid, not the bench): apply and CG rates per configuration."""
import argparse
import json
import sys
import os

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import thirring2d_b200 as tb

PEAK = 6550.7


def probe(nt, nx, C, m, rows=0, chunk=0, solver=0, reps=20, max_iter=100000):
    dev = torch.device("cuda:0")
    rng = np.random.default_rng(1)
    stream = torch.cuda.current_stream()
    ctx = tb.Context(nt, nx, C, tb.MODE_ADJOINT, m=m, mu=0.0, stream=stream.cuda_stream)
    ctx.set_tuning(rows, chunk, solver)
    ctx.set_cg(1e-30, max_iter)
    A = torch.from_numpy(rng.uniform(-np.pi, np.pi, size=(C, nt, nx, 2))).to(dev)
    ctx.set_gauge_dev(A.data_ptr())
    n = ctx.vec_doubles
    v = torch.randn(n, dtype=torch.float64, device=dev)
    o = torch.empty_like(v)
    b = torch.empty_like(v)
    x = torch.empty_like(v)
    for _ in range(3):
        ctx.apply_dev(tb.OP_M, v.data_ptr(), o.data_ptr())
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        ctx.apply_dev(tb.OP_M, v.data_ptr(), o.data_ptr())
        ctx.apply_dev(tb.OP_MDAG, o.data_ptr(), b.data_ptr())
    e1.record()
    torch.cuda.synchronize()
    ms_apply = e0.elapsed_time(e1) / (2 * reps)
    sites = nt * nx * C
    out = {"cfg": f"{nt}x{nx}x{C} m={m} rows={rows} chunk={chunk} solver={solver}",
           "apply_us": round(ms_apply * 1e3, 2),
           "apply_GBs": round(64 * sites / ms_apply / 1e6, 1),
           "apply_frac": round(64 * sites / ms_apply / 1e6 / PEAK, 3)}
    ctx.apply_dev(tb.OP_MDAG, v.data_ptr(), b.data_ptr())
    ctx.cg_dev(b.data_ptr(), x.data_ptr())  # warm (graph build)
    ctx.cg_dev(b.data_ptr(), x.data_ptr())
    info = ctx.cg_result()
    ms = ctx.last_solve_ms
    its = float(info.iters.mean())
    imax = int(info.iters.max())
    isum = float(info.iters.astype(np.int64).sum())
    out.update({"cg_ms": round(ms, 3), "iters_mean": its, "iters_max": imax,
                "us_per_iter": round(ms * 1e3 / max(imax, 1), 2),
                "cg_GBs": round(288.0 * nt * nx * isum / ms / 1e6, 1),
                "cg_frac": round(288.0 * nt * nx * isum / ms / 1e6 / PEAK, 3),
                "status": np.bincount(info.status, minlength=4).tolist()})
    ctx.close()
    return out


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--cfg", nargs="*", default=None)
    ap.add_argument("--one", nargs=3, type=int, default=None, help="NT NX C: one streaming-solver run, 60 iterations")
    args = ap.parse_args()
    if args.one:
        print(json.dumps(probe(args.one[0], args.one[1], args.one[2], 0.01, solver=1, reps=3, max_iter=61)))
        sys.exit(0)
    cfgs = [(64, 64, 256, 0.1), (64, 64, 1024, 0.1), (256, 256, 64, 0.1), (2048, 2048, 1, 0.1), (32, 32, 1, 1.0),
            (128, 128, 512, 0.1)]
    for c in cfgs:
        mi = 100000 if c[0] * c[1] * c[2] < 3e6 else 400
        for rows in (0,):
            print(json.dumps(probe(*c, rows=rows, max_iter=mi)), flush=True)
