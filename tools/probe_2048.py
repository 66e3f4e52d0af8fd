"""2048^2 single lattice, streaming solver, 200 iterations (development aid): which kernels, how fast."""
import os, sys, json
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import thirring2d_b200 as tb
dev = torch.device("cuda:0")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
for m in (0.05, 0.01):
    gen = torch.Generator(device=dev).manual_seed(5)
    A = ((torch.rand(n * n * 2, dtype=torch.float64, device=dev, generator=gen) - 0.5) * (2 * np.pi)).view(n, n, 2)
    b = torch.randn(n * n * 2, dtype=torch.float64, device=dev, generator=gen).view(n, n, 2)
    ctx = tb.Context(n, n, 1, tb.MODE_ADJOINT, device=0, m=m, mu=0.0, stream=torch.cuda.current_stream().cuda_stream)
    ctx.set_cg(1e-30, 201)
    ctx.set_tuning(0, 0, 1)
    ctx.set_gauge_dev(A.data_ptr())
    x = torch.empty_like(b)
    ms = []
    for _ in range(4):
        ctx.cg_dev(b.data_ptr(), x.data_ptr())
        ms.append(ctx.last_solve_ms)
    it = int(ctx.cg_result().iters.max())
    print(json.dumps({"n": n, "m": m, "kernels": ctx.streaming_info(), "us_per_iter": [round(v * 1e3 / it, 2) for v in ms], "iters": it}))
    ctx.close()
