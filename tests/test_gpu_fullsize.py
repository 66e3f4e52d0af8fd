"""Parity at BASELINE.json's full sizes through size-independent properties (the oracle would need minutes to
hours there): adjointness, linearity, the CG residual checked by an independent apply, agreement of the two
solvers, and a checksum of the batched result against single-chain solves of a few sampled chains."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")
tb = pytest.importorskip("thirring2d_b200")


def dev_vec(ctx, gen):
    return torch.randn(ctx.vec_doubles, dtype=torch.float64, device="cuda", generator=gen)


def cdot(a, b):
    """<a,b> per whole batch for interleaved (re,im) device vectors."""
    ar, ai, br, bi = a[0::2], a[1::2], b[0::2], b[1::2]
    return complex(float((ar * br + ai * bi).sum()), float((ar * bi - ai * br).sum()))


@pytest.mark.parametrize("nt,nx,nchains", [(64, 64, 256), (256, 256, 8), (2048, 2048, 1), (128, 128, 64)])
def test_apply_adjointness_and_linearity_full_size(nt, nx, nchains):
    gen = torch.Generator(device="cuda").manual_seed(nt + nchains)
    with tb.Context(nt, nx, nchains, tb.MODE_ADJOINT, m=0.05, mu=0.1,
                    stream=torch.cuda.current_stream().cuda_stream) as ctx:
        A = (torch.rand(nchains * nt * nx * 2, dtype=torch.float64, device="cuda", generator=gen) - 0.5) * (2 * np.pi)
        ctx.set_gauge_dev(A.data_ptr())
        a, b = dev_vec(ctx, gen), dev_vec(ctx, gen)
        Mb, Mda, Mab = torch.empty_like(a), torch.empty_like(a), torch.empty_like(a)
        ctx.apply_dev(tb.OP_M, b.data_ptr(), Mb.data_ptr())
        ctx.apply_dev(tb.OP_MDAG, a.data_ptr(), Mda.data_ptr())
        lhs, rhs = cdot(a, Mb), cdot(Mda, b)
        assert abs(lhs - rhs) <= 1e-12 * abs(lhs) + 1e-7, (lhs, rhs)       # <a, M b> = <M^dagger a, b>
        # linearity: M(a + 2 b) = M a + 2 M b
        ab = a + 2.0 * b
        Ma = torch.empty_like(a)
        ctx.apply_dev(tb.OP_M, a.data_ptr(), Ma.data_ptr())
        ctx.apply_dev(tb.OP_M, ab.data_ptr(), Mab.data_ptr())
        err = float((Mab - (Ma + 2.0 * Mb)).norm() / Mab.norm())
        assert err <= 1e-14
        # M~M is Hermitian positive: <b, M^dagger M b> = |M b|^2
        MdMb = torch.empty_like(a)
        ctx.apply_dev(tb.OP_MDM, b.data_ptr(), MdMb.data_ptr())
        q = cdot(b, MdMb)
        assert abs(q.imag) <= 1e-12 * q.real and abs(q.real - float((Mb * Mb).sum())) <= 1e-12 * q.real


@pytest.mark.parametrize("nt,nx,nchains,m,solvers", [(64, 64, 256, 0.1, (2, 1)), (256, 256, 8, 0.05, (1,)),
                                                      (1024, 1024, 1, 0.1, (1,))])
def test_cg_residual_full_size(nt, nx, nchains, m, solvers):
    """The true residual |b - M~M x| / |b| is at rounding level, checked with applies that are independent of
    the solver's recursion; where both solvers exist they agree to 1e-12 and in iteration count +-1."""
    gen = torch.Generator(device="cuda").manual_seed(5)
    sols, iters = [], []
    with tb.Context(nt, nx, nchains, tb.MODE_ADJOINT, m=m, mu=0.0,
                    stream=torch.cuda.current_stream().cuda_stream) as ctx:
        A = torch.randn(nchains * nt * nx * 2, dtype=torch.float64, device="cuda", generator=gen) * 0.6
        ctx.set_gauge_dev(A.data_ptr())
        xi = dev_vec(ctx, gen)
        b, x, chk = torch.empty_like(xi), torch.empty_like(xi), torch.empty_like(xi)
        ctx.apply_dev(tb.OP_MCONJ, xi.data_ptr(), b.data_ptr())
        for solver in solvers:
            ctx.set_tuning(solver=solver)
            ctx.cg_dev(b.data_ptr(), x.data_ptr())
            info = ctx.cg_result()
            assert np.all(info.status == tb.CG_CONVERGED)
            ctx.apply_dev(tb.OP_MDM, x.data_ptr(), chk.data_ptr())
            res = float((chk - b).norm() / b.norm())
            assert res <= 5e-12, res
            sols.append(x.clone())
            iters.append(info.iters.copy())
    if len(sols) == 2:
        assert float((sols[0] - sols[1]).norm() / sols[1].norm()) <= 1e-12
        assert np.all(np.abs(iters[0].astype(int) - iters[1].astype(int)) <= 1)


def test_batched_solution_equals_single_chain_solutions():
    """Chains are independent: chain c of the 256-chain batch equals the same chain solved alone (sampled)."""
    nt = nx = 64
    n = 256
    rng = np.random.default_rng(0)
    A = rng.vonmises(0.0, 2 / 0.3, size=(n, nt, nx, 2))
    xi = rng.normal(size=(n, nt, nx)) + 1j * rng.normal(size=(n, nt, nx))
    with tb.Context(nt, nx, n, tb.MODE_ADJOINT, m=0.1) as ctx:
        ctx.set_gauge(A)
        b = ctx.fm_conjugate_mul(xi)
        x, info = ctx.fmdm_invert_cg(b)
    for c in (0, 101, 255):
        with tb.Context(nt, nx, 1, tb.MODE_ADJOINT, m=0.1) as one:
            one.set_gauge(A[c])
            x1, i1 = one.fmdm_invert_cg(b[c])
        assert np.array_equal(x1, x[c])          # same kernel, same per-chain arithmetic: bitwise
        assert i1.iters[0] == info.iters[c]


@pytest.mark.parametrize("nt,nx,nchains", [(256, 256, 2), (2048, 2048, 1)])
def test_T1_apply_matches_oracle_at_full_size(oracle, nt, nx, nchains):
    """T1 of SURVEY 8(c) at the large lattices of BASELINE.json directly against the C oracle (a 2048^2 apply takes it
    under a second): M, M^dagger and M~ to 1e-13 in l2 and max norm, per-chain mu."""
    from tests.util import APPLY_TOL, assert_close, random_gauge, random_vector

    rng = np.random.default_rng(nt + nchains)
    A = random_gauge(rng, nchains, nt, nx)
    v = random_vector(rng, nchains, nt, nx)
    m = rng.uniform(0.05, 1.0, size=nchains)
    mu = rng.uniform(-0.2, 0.2, size=nchains)
    with tb.Context(nt, nx, nchains, tb.MODE_ADJOINT) as ctx:
        ctx.set_params(m, mu)
        ctx.set_gauge(A)
        got_m, got_d = ctx.fm_mul(v), ctx.fm_conjugate_mul(v)
    for c in range(nchains):
        assert_close(got_m[c], oracle.fm_mul(v[c], A[c], m[c], mu[c]), APPLY_TOL, "M")
        assert_close(got_d[c], oracle.fm_conjugate_mul(v[c], A[c], m[c], mu[c], tb.MODE_ADJOINT), APPLY_TOL, "M^dagger")


@pytest.mark.parametrize("n,mode,mu", [(200, tb.MODE_ADJOINT, 0.0), (3, tb.MODE_ADJOINT, 0.1), (5, tb.MODE_REF_COMPAT, 0.05)])
def test_host_buffer_path_through_the_canonical_layout_kernel(oracle, n, mode, mu):
    """tb_cg and tb_cg_gauge on a 64^2 context go host buffer -> H2D -> ONE kernel (canonical layout in and out, links
    built inside from the angles) -> D2H per sub-batch of chains.  The result must be bitwise the device-resident
    path's (pack -> links kernel -> solver -> unpack), the context must be left as tb_set_gauge leaves it (links and
    angles), and a sample of chains is checked against the oracle.  200 chains: two waves on 148 SMs (4 + 4 sub-batches)."""
    import torch

    nt = nx = 64
    m = 100.0 if mode == tb.MODE_REF_COMPAT else 0.2
    rng = np.random.default_rng(n)
    A = rng.uniform(-np.pi, np.pi, size=(n, nt, nx, 2))
    A2 = rng.uniform(-np.pi, np.pi, size=(n, nt, nx, 2))
    xi = rng.normal(size=(n, nt, nx)) + 1j * rng.normal(size=(n, nt, nx))
    dev = torch.device("cuda", 0)
    with tb.Context(nt, nx, n, mode, m=m, mu=mu) as ctx:
        # device-resident reference: explicit re-layout kernels and links_kernel
        A_dev = torch.from_numpy(A).to(dev)
        ctx.set_gauge_dev(A_dev.data_ptr())
        v_canon = torch.from_numpy(xi.view(np.float64)).to(dev)
        v, b, x = (torch.empty(ctx.vec_doubles, dtype=torch.float64, device=dev) for _ in range(3))
        ctx.pack_dev(v_canon.data_ptr(), v.data_ptr())
        ctx.apply_dev(tb.OP_MCONJ, v.data_ptr(), b.data_ptr())
        os.environ["TB_NO_PLAN"] = "1"
        ctx.cg_dev(b.data_ptr(), x.data_ptr())
        del os.environ["TB_NO_PLAN"]
        it_dev = ctx.cg_result().iters.copy()
        out = torch.empty_like(v_canon)
        ctx.unpack_dev(b.data_ptr(), out.data_ptr())
        b_host = out.cpu().numpy().view(np.complex128).reshape(n, nt, nx)
        ctx.unpack_dev(x.data_ptr(), out.data_ptr())
        x_dev = out.cpu().numpy().view(np.complex128).reshape(n, nt, nx)
        # host-buffer entry points
        ctx.set_gauge(A2)                                  # something else, so that cg_gauge has to install A
        x1, i1 = ctx.fmdm_invert_cg_with_gauge(A, b_host)   # tb_cg_gauge: links built inside the solver kernel
        assert np.array_equal(x1, x_dev) and np.array_equal(i1.iters, it_dev)
        assert np.array_equal(ctx.get_gauge(), A)           # the context's angles ...
        assert np.array_equal(ctx.fm_conjugate_mul(xi), b_host)   # ... and links are the new field's
        x2, i2 = ctx.fmdm_invert_cg(b_host)                 # tb_cg: links from the context
        assert np.array_equal(x2, x_dev) and np.array_equal(i2.iters, it_dev)
        os.environ["TB_NO_CANON"] = "1"                     # the re-layout path stays available and agrees
        x3, i3 = ctx.fmdm_invert_cg_with_gauge(A, b_host)
        del os.environ["TB_NO_CANON"]
        assert np.array_equal(x3, x_dev)
    for c in sorted({0, n // 2, n - 1}):
        xo, st, it, rr = oracle.fmdm_invert_cg(b_host[c], A[c], m, mu, mode)
        assert abs(int(it_dev[c]) - it) <= 1
        assert np.linalg.norm(x1[c] - xo) <= 1e-12 * np.linalg.norm(xo)
