"""Small solves for compute-sanitizer (memcheck / racecheck / synccheck): every on-chip kernel shape once, few chains."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import thirring2d_b200 as tb

rng = np.random.default_rng(0)
only = sys.argv[1] if len(sys.argv) > 1 else ""   # 'small' = one-CTA kernels, 'cluster' = cluster kernels
CASES = [(16, 16, 3, tb.MODE_ADJOINT, 0.5, 0.1), (32, 32, 2, tb.MODE_ADJOINT, 0.5, 0.0),
                                 (64, 64, 2, tb.MODE_ADJOINT, 0.5, 0.1), (64, 64, 1, tb.MODE_REF_COMPAT, 50.0, 0.1),
                                 (128, 128, 1, tb.MODE_ADJOINT, 0.6, 0.1), (128, 64, 1, tb.MODE_REF_COMPAT, 50.0, 0.0),
                                 (256, 256, 1, tb.MODE_ADJOINT, 1.0, 0.0), (48, 40, 2, tb.MODE_ADJOINT, 0.7, 0.1)]
if only in ("new", "pipe"):
    # planned launches (a second solve: the first one provides the iteration estimates) and the TMA-staged kernels
    for (nt, nx, n, solver) in [(64, 64, 160, 0), (128, 128, 40, 0), (256, 256, 8, 0), (64, 128, 8, 1), (32, 32, 64, 1)]:
        if only == "pipe" and solver != 1:
            continue
        A = rng.uniform(-np.pi, np.pi, size=(n, nt, nx, 2))
        xi = rng.normal(size=(n, nt, nx)) + 1j * rng.normal(size=(n, nt, nx))
        os.environ["TB_SUBBATCHES"] = "1"
        os.environ["TB_PIPE_TEST"] = "1"
        with tb.Context(nt, nx, n, tb.MODE_ADJOINT, m=np.linspace(1.5, 3.0, n), mu=0.1) as ctx:
            ctx.set_tuning(solver=solver)
            ctx.set_gauge(A)
            b = ctx.fm_conjugate_mul(xi)
            for rep in range(2):
                x, info = ctx.fmdm_invert_cg(b)
            print(nt, nx, n, ctx.solver_info(), int(info.iters.min()), int(info.iters.max()),
                  np.bincount(info.status, minlength=4).tolist(), flush=True)
    sys.exit(0)
if only == "r2":
    # round 2: 16-chain tiles (one ragged), the two-launch iteration, a shared gauge field, the point-to-point cluster
    # hand-overs; few iterations (memcheck)
    os.environ["TB_SUBBATCHES"] = "1"
    os.environ["TB_PIPE_TEST"] = "1"
    for (nt, nx, n, xpay, shared) in [(32, 32, 64, "0", False), (32, 32, 20, "1", False), (32, 64, 8, "1", True),
                                      (32, 32, 40, "0", True), (16, 48, 48, "1", True)]:
        os.environ["TB_PIPE_XPAY"] = xpay
        A = rng.uniform(-np.pi, np.pi, size=(n, nt, nx, 2))
        xi = rng.normal(size=(n, nt, nx)) + 1j * rng.normal(size=(n, nt, nx))
        with tb.Context(nt, nx, n, tb.MODE_ADJOINT, m=0.5, mu=0.1) as ctx:
            ctx.set_tuning(solver=1)
            ctx.set_cg(1e-30, 12)
            if shared:
                ctx.set_gauge_shared(A[0])
            else:
                ctx.set_gauge(A)
            b = ctx.fm_conjugate_mul(xi)
            x, info = ctx.fmdm_invert_cg(b)
            print(nt, nx, n, "xpay", xpay, "shared", shared, ctx.streaming_info(), int(info.iters.max()), flush=True)
    for (nt, nx) in [(128, 128), (256, 256), (128, 64)]:
        A = rng.uniform(-np.pi, np.pi, size=(2, nt, nx, 2))
        xi = rng.normal(size=(2, nt, nx)) + 1j * rng.normal(size=(2, nt, nx))
        with tb.Context(nt, nx, 2, tb.MODE_ADJOINT, m=0.5, mu=0.1) as ctx:
            ctx.set_cg(1e-30, 12)
            ctx.set_gauge(A)
            b = ctx.fm_conjugate_mul(xi)
            x, info = ctx.fmdm_invert_cg(b)
            print(nt, nx, ctx.solver_info(), info.iters.tolist(), flush=True)
    sys.exit(0)
for (nt, nx, n, mode, m, mu) in CASES:
    if (only == "small" and nt * nx > 4096) or (only == "cluster" and nt * nx <= 4096):
        continue
    A = rng.uniform(-np.pi, np.pi, size=(n, nt, nx, 2))
    xi = rng.normal(size=(n, nt, nx)) + 1j * rng.normal(size=(n, nt, nx))
    with tb.Context(nt, nx, n, mode, m=m, mu=mu) as ctx:
        ctx.set_cg(1e-30, 40)
        ctx.set_gauge(A)
        b = ctx.fm_conjugate_mul(xi)
        x, info = ctx.fmdm_invert_cg(b)
        print(nt, nx, ctx.solver_info(), info.iters.tolist(), info.status.tolist(), flush=True)
# family B on the masked on-chip kernel
if only == "cluster":
    sys.exit(0)
field = (rng.random((2, 64, 64)) < 0.1).astype(np.int32)
with tb.Context(64, 64, 2, tb.MODE_ADJOINT, m=0.3, mu=0.1) as ctx:
    ctx.set_cg(1e-30, 30)
    ctx.set_occupancy(field)
    x, info = ctx.cg_MdM(rng.normal(size=(2, 64, 64)))
    print("family B", ctx.solver_info(), info.iters.tolist(), flush=True)
