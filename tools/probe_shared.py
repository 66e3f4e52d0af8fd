"""Multi-RHS solves on ONE gauge field (development aid): the field uploaded once per chain (tb_set_gauge_dev) against
tb_set_gauge_shared_dev (compact links in the staged streaming kernels).  python tools/probe_shared.py [NT,NX,C ...]"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import thirring2d_b200 as tb


def run(nt, nx, C, shared, iters=200):
    dev = torch.device("cuda:0")
    rng = np.random.default_rng(1)
    ctx = tb.Context(nt, nx, C, tb.MODE_ADJOINT, m=0.01, mu=0.0, stream=torch.cuda.current_stream().cuda_stream)
    ctx.set_tuning(0, 0, 1)
    ctx.set_cg(1e-30, iters + 1)
    A1 = torch.from_numpy(rng.uniform(-np.pi, np.pi, size=(nt, nx, 2))).to(dev)
    if shared:
        ctx.set_gauge_shared_dev(A1.data_ptr())
    else:
        A = A1.unsqueeze(0).expand(C, nt, nx, 2).contiguous()
        ctx.set_gauge_dev(A.data_ptr())
    n = ctx.vec_doubles
    b = torch.randn(n, dtype=torch.float64, device=dev)
    x = torch.empty_like(b)
    v = torch.empty_like(b)
    ctx.cg_dev(b.data_ptr(), x.data_ptr())
    ms = []
    for _ in range(3):
        ctx.cg_dev(b.data_ptr(), x.data_ptr())
        ms.append(ctx.last_solve_ms)
    it = int(ctx.cg_result().iters.max())
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(2):
        ctx.apply_dev(tb.OP_M, b.data_ptr(), v.data_ptr())
    torch.cuda.synchronize()
    e0.record()
    for _ in range(10):
        ctx.apply_dev(tb.OP_M, b.data_ptr(), v.data_ptr())
        ctx.apply_dev(tb.OP_MDAG, v.data_ptr(), x.data_ptr())
    e1.record()
    torch.cuda.synchronize()
    out = {"cfg": f"{nt}x{nx} x {C} sources", "shared": shared, "kernels": ctx.streaming_info(),
           "us_per_iter": round(min(ms) * 1e3 / it, 2), "apply_us": round(e0.elapsed_time(e1) * 1e3 / 20, 2)}
    ctx.close()
    return out


if __name__ == "__main__":
    cfgs = [(512, 512, 32), (1024, 1024, 16), (256, 256, 64), (512, 512, 20), (2048, 2048, 4)]
    if len(sys.argv) > 1:
        cfgs = [tuple(int(v) for v in a.split(",")) for a in sys.argv[1:]]
    for c in cfgs:
        for shared in (False, True):
            print(json.dumps(run(*c, shared)), flush=True)
