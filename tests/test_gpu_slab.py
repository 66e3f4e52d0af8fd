"""T7: slab decomposition over 2, 4 and 8 GPUs against the single-GPU path (needs that many GPUs; `gpurun --gpus N`).
One process per GPU; halos are peer reads over NVLink, dot products one-shot peer all-reduces."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")
tb = pytest.importorskip("thirring2d_b200")


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _inputs(nt, nx, nchains, seed):
    rng = np.random.default_rng(seed)
    A = rng.uniform(-np.pi, np.pi, size=(nchains, nt, nx, 2))
    v = rng.normal(size=(nchains, nt, nx)) + 1j * rng.normal(size=(nchains, nt, nx))
    return A, v


def _worker(rank, world, port, nt, nx, nchains, m, mu, q):
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    A, v = _inputs(nt, nx, nchains, 77)
    ntl = nt // world
    sl = slice(rank * ntl, (rank + 1) * ntl)
    ctx = tb.Context(nt, nx, nchains, tb.MODE_ADJOINT, device=rank, m=m, mu=mu, slab_rank=rank, slab_nranks=world)
    ctx.slab_setup(dist)
    dist.barrier()
    ctx.set_gauge(A[:, sl])
    ctx.synchronize()
    dist.barrier()
    out = {}
    for name, op in (("M", tb.OP_M), ("Mdag", tb.OP_MDAG), ("MdM", tb.OP_MDM)):
        out[name] = ctx.apply(op, v[:, sl])
    b = ctx.fm_conjugate_mul(v[:, sl])
    x, info = ctx.fmdm_invert_cg(b)
    x2, info2 = ctx.fmdm_invert_cg(b)  # a second solve exercises the epoch hand-over between solves
    xi, infoi = ctx.fm_invert_cg(v[:, sl])
    ctx.set_tuning(solver=3)   # the 4-kernel iteration with separate scalar kernels (the form REF_COMPAT uses)
    x4, info4 = ctx.fmdm_invert_cg(b)
    # the fused iteration as separate kernels (TB_NO_PERSIST), then the one-launch solve again: both protocols share
    # the epoch counter and the neighbour flags
    ctx.set_tuning(solver=0)
    os.environ["TB_NO_PERSIST"] = "1"
    xm, infom = ctx.fmdm_invert_cg(b)
    del os.environ["TB_NO_PERSIST"]
    x5, info5 = ctx.fmdm_invert_cg(b)
    # every form of the one-launch solve explicitly (the default picks one by slab size): block 0 runs the all-reduce
    # (TB_SLAB_SYNC=0); the last-arriving block exchanges with the peers and publishes the totals (1); every block polls
    # the peers' slots (2), here with a single slot replica; then the default again
    os.environ["TB_SLAB_SYNC"] = "0"
    x6, info6 = ctx.fmdm_invert_cg(b)
    os.environ["TB_SLAB_SYNC"] = "1"
    x9, info9 = ctx.fmdm_invert_cg(b)
    os.environ["TB_SLAB_SYNC"] = "2"
    os.environ["TB_SLAB_NREP"] = "1"
    x7, info7 = ctx.fmdm_invert_cg(b)
    del os.environ["TB_SLAB_NREP"], os.environ["TB_SLAB_SYNC"]
    x8, info8 = ctx.fmdm_invert_cg(b)
    os.environ["TB_SLAB_SYSFENCE"] = "1"   # system-scope instead of device-scope fences around the peer exchange
    xf, infof = ctx.fmdm_invert_cg(b)
    del os.environ["TB_SLAB_SYSFENCE"]
    out.update(b=b, x=x, x2=x2, xi=xi, x4=x4, iters=info.iters, status=info.status, iters2=info2.iters,
               iters4=info4.iters, xm=xm, itersm=infom.iters, x5=x5, iters5=info5.iters, x6=x6, iters6=info6.iters,
               x7=x7, iters7=info7.iters, x8=x8, iters8=info8.iters, x9=x9, iters9=info9.iters, xf=xf, itersf=infof.iters)
    q.put((rank, out))
    dist.barrier()
    ctx.close()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4, 8], ids=["P2", "P4", "P8"])
@pytest.mark.parametrize("nt,nx,nchains,m,mu", [(32, 32, 1, 0.2, 0.1), (64, 48, 3, 0.1, 0.0), (16, 16, 40, 0.5, 0.2), (256, 512, 1, 0.1, 0.0)])
def test_T7_slab_matches_single_gpu(nt, nx, nchains, m, mu, world):
    """Rank-order sums, the wrap link between rank P-1 and rank 0, and both forms of the one-launch solve at every
    rank count the box offers (hmc.c:364-395 is the loop all of them restate)."""
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    import torch.multiprocessing as mp

    mpctx = mp.get_context("spawn")
    q = mpctx.Queue()
    port = _free_port()
    procs = [mpctx.Process(target=_worker, args=(r, world, port, nt, nx, nchains, m, mu, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=300) for _ in range(world))
    for p in procs:
        p.join(timeout=300)
        assert p.exitcode == 0

    A, v = _inputs(nt, nx, nchains, 77)
    with tb.Context(nt, nx, nchains, tb.MODE_ADJOINT, device=0, m=m, mu=mu) as ctx:
        ctx.set_tuning(solver=1)  # the streaming solver: same per-site arithmetic as the slab kernels
        ctx.set_gauge(A)
        ref = {"M": ctx.apply(tb.OP_M, v), "Mdag": ctx.apply(tb.OP_MDAG, v), "MdM": ctx.apply(tb.OP_MDM, v)}
        b = ctx.fm_conjugate_mul(v)
        x, info = ctx.fmdm_invert_cg(b)
        xi, infoi = ctx.fm_invert_cg(v)
    cat = lambda k: np.concatenate([res[r][k] for r in range(world)], axis=1)
    for k in ("M", "Mdag", "MdM"):
        assert np.array_equal(cat(k), ref[k]), f"slab apply {k} is not bitwise identical to the single-GPU apply"
    assert np.array_equal(cat("b"), b)
    for r in range(world):
        assert np.all(res[r]["status"] == tb.CG_CONVERGED)
        assert np.array_equal(res[r]["iters"], res[0]["iters"])  # identical decisions on every rank
        assert np.all(np.abs(res[r]["iters"].astype(int) - info.iters.astype(int)) <= 1)
        assert np.array_equal(res[r]["iters"], res[r]["iters2"])
    xs = cat("x")
    assert np.array_equal(xs, cat("x2"))
    assert np.linalg.norm(cat("x4") - x) <= 1e-12 * np.linalg.norm(x)
    assert np.all(np.abs(res[0]["iters4"].astype(int) - info.iters.astype(int)) <= 1)
    assert np.linalg.norm(xs - x) <= 1e-12 * np.linalg.norm(x)
    assert np.linalg.norm(cat("xi") - xi) <= 1e-12 * np.linalg.norm(xi)
    assert np.linalg.norm(cat("xm") - x) <= 1e-12 * np.linalg.norm(x)
    assert np.all(np.abs(res[0]["itersm"].astype(int) - info.iters.astype(int)) <= 1)
    assert np.array_equal(cat("x5"), xs) and np.array_equal(res[0]["iters5"], res[0]["iters"])
    # block 0's form and the other two cut the slab differently (another summation order of the partials): the same
    # solve to rounding, identical decisions on every rank
    for k in ("6", "9"):
        assert np.linalg.norm(cat("x" + k) - x) <= 1e-12 * np.linalg.norm(x)
        assert np.all(np.abs(res[0]["iters" + k].astype(int) - info.iters.astype(int)) <= 1)
        for r in range(world):
            assert np.array_equal(res[r]["iters" + k], res[0]["iters" + k])
    # who polls does not change a bit of the arithmetic
    assert np.array_equal(cat("x7"), cat("x9")) and np.array_equal(res[0]["iters7"], res[0]["iters9"])
    assert np.array_equal(cat("x8"), xs) and np.array_equal(res[0]["iters8"], res[0]["iters"])
    assert np.array_equal(cat("xf"), xs) and np.array_equal(res[0]["itersf"], res[0]["iters"])   # the scope of a fence changes no bit
