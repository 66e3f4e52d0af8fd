"""Planned launches of the 64x64 on-chip solver (tb_resident.cu: TbPlan): a batch larger than the SM count is cut
into one equal share of CG iterations per SM, a chain that straddles two shares is paused on one SM and resumed on
another.  The arithmetic of a chain must not notice: x, the iteration count, the status and the final residual are
BITWISE those of the one-CTA-per-chain launch (TB_NO_PLAN=1), whatever the quality of the iteration estimates."""
import os

import numpy as np
import pytest

from tests.util import CG_SOL_TOL, assert_close, random_vector, smooth_gauge

pytestmark = pytest.mark.gpu

tb = pytest.importorskip("thirring2d_b200")

NT = NX = 64


@pytest.fixture
def whole_batch(monkeypatch):
    monkeypatch.setenv("TB_SUBBATCHES", "1")   # the host path then solves the batch in one launch, like tb_cg_dev
    monkeypatch.delenv("TB_NO_PLAN", raising=False)


def solve(ctx, b, planned):
    if planned:
        os.environ.pop("TB_NO_PLAN", None)
    else:
        os.environ["TB_NO_PLAN"] = "1"
    try:
        return ctx.fmdm_invert_cg(b)
    finally:
        os.environ.pop("TB_NO_PLAN", None)


def same(a, b):
    (xa, ia), (xb, ib) = a, b
    assert np.array_equal(ia.status, ib.status)
    assert np.array_equal(ia.iters, ib.iters)
    assert np.array_equal(ia.rr, ib.rr)
    assert np.array_equal(xa, xb)


@pytest.mark.parametrize("mode,mu", [(tb.MODE_ADJOINT, 0.0), (tb.MODE_ADJOINT, 0.1), (tb.MODE_REF_COMPAT, 0.05)])
def test_planned_launch_is_bitwise_the_plain_launch(whole_batch, oracle, mode, mu):
    C = 211
    rng = np.random.default_rng(17)
    A = smooth_gauge(rng, C, NT, NX, 0.4)
    xi = random_vector(rng, C, NT, NX)
    xi[5] = 0.0                                        # a zero source among them (hmc.c:359-361)
    # ragged masses: 40 ... 400 iterations per chain (REF_COMPAT, M~ = M, needs a heavy mass to converge at all)
    m = np.exp(rng.uniform(np.log(0.05), np.log(1.0), size=C)) if mode == tb.MODE_ADJOINT else rng.uniform(60, 200, size=C)
    with tb.Context(NT, NX, C, mode, m=m, mu=mu) as ctx:
        ctx.set_gauge(A)
        if mode == tb.MODE_REF_COMPAT:
            # M.M is not Hermitian: whether the recursive residual ever falls below the reference's absolute 1e-30
            # from ||b||^2 ~ 1e8 is a matter of luck, so this case stops at 1e-6 (1e-14 relative)
            ctx.set_cg(1e-6, 100000)
        b = ctx.fm_conjugate_mul(xi)
        ref = solve(ctx, b, planned=False)
        assert ref[1].status[5] == tb.CG_ZERO_SOURCE and np.all(np.delete(ref[1].status, 5) == tb.CG_CONVERGED)
        first = solve(ctx, b, planned=True)            # chain 5 did not "converge": chains are dealt out whole
        same(first, ref)
        # estimates of every chain valid from here on: shares with split chains
        b2 = b.copy()
        b2[5] = b[6]
        ref2 = solve(ctx, b2, planned=False)
        same(solve(ctx, b2, planned=True), ref2)
        same(solve(ctx, b2, planned=True), ref2)
        # estimates that are badly wrong: the masses (hence the iteration counts) change between the solves
        ctx.set_params(m[::-1].copy(), mu)
        ref3 = solve(ctx, b2, planned=False)
        ctx.set_params(m, mu)
        solve(ctx, b2, planned=False)                  # estimates of the forward order ...
        ctx.set_params(m[::-1].copy(), mu)
        same(solve(ctx, b2, planned=True), ref3)       # ... used for the reversed one
        # and against the oracle, two chains
        ctx.set_params(m, mu)
        x, info = solve(ctx, b2, planned=True)
        for c in (0, C - 1) if mode == tb.MODE_ADJOINT else ():
            xo, st, it, rr = oracle.fmdm_invert_cg(b2[c], A[c], float(m[c]), mu, mode)
            assert st == info.status[c] and abs(it - int(info.iters[c])) <= 1
            assert_close(x[c], xo, CG_SOL_TOL, "planned launch vs oracle")


def test_planned_launch_with_max_iter_inside_a_head(whole_batch):
    """Estimates say ~280 iterations, then max_iter drops to 60: the chains end (status max-iter) inside their heads
    and the CTAs that hold the tails must skip them."""
    C = 256
    rng = np.random.default_rng(3)
    A = smooth_gauge(rng, C, NT, NX, 0.5)
    xi = random_vector(rng, C, NT, NX)
    with tb.Context(NT, NX, C, tb.MODE_ADJOINT, m=0.1, mu=0.0) as ctx:
        ctx.set_gauge(A)
        b = ctx.fm_conjugate_mul(xi)
        full = solve(ctx, b, planned=False)
        assert np.all(full[1].status == tb.CG_CONVERGED) and full[1].iters.min() > 100
        same(solve(ctx, b, planned=True), full)
        ctx.set_cg(1e-30, 61)
        short = solve(ctx, b, planned=True)            # planned with the long estimates
        assert np.all(short[1].status == tb.CG_MAXITER) and np.all(short[1].iters == 60)
        same(solve(ctx, b, planned=False), short)
        same(solve(ctx, b, planned=True), short)       # previous solve not converged: dealt out whole


@pytest.mark.parametrize("nt,nx,C,mu", [(128, 128, 40, 0.0), (128, 128, 50, 0.1), (256, 256, 8, 0.0), (64, 128, 90, 0.0)])
def test_planned_cluster_launch_is_bitwise_the_plain_launch(whole_batch, oracle, nt, nx, C, mu):
    """The same for the cluster solver (tb_cluster.cu): a batch that fills its last wave of clusters badly (8 chains of
    256^2 on 7 co-resident clusters of 16 CTAs) is cut into one share per cluster; the state of a split chain (r, p, x
    of every slab) crosses from one cluster to another through HBM."""
    rng = np.random.default_rng(29)
    A = smooth_gauge(rng, C, nt, nx, 0.4)
    xi = random_vector(rng, C, nt, nx)
    xi[1] = 0.0
    m = np.exp(rng.uniform(np.log(0.05), np.log(0.6), size=C))
    with tb.Context(nt, nx, C, tb.MODE_ADJOINT, m=m, mu=mu) as ctx:
        kind, capacity = ctx.solver_info()
        if kind != 2 or C <= capacity:
            pytest.skip(f"no cluster shape on this device, or one wave holds the batch: kind {kind}, {capacity} clusters")
        ctx.set_gauge(A)
        b = ctx.fm_conjugate_mul(xi)
        n0 = ctx.launch_count
        ref = solve(ctx, b, planned=False)
        n1 = ctx.launch_count
        solve(ctx, b, planned=True)
        assert ctx.launch_count - n1 == n1 - n0 + 1, "the planned path (one more launch: the scheduler) did not run"
        assert ref[1].status[1] == tb.CG_ZERO_SOURCE and np.all(np.delete(ref[1].status, 1) == tb.CG_CONVERGED)
        same(solve(ctx, b, planned=True), ref)          # a chain without an estimate: dealt out whole
        b2 = b.copy()
        b2[1] = b[2]
        ref2 = solve(ctx, b2, planned=False)
        same(solve(ctx, b2, planned=True), ref2)        # split chains
        same(solve(ctx, b2, planned=True), ref2)
        ctx.set_params(m[::-1].copy(), mu)              # estimates that are badly wrong
        ref3 = solve(ctx, b2, planned=False)
        ctx.set_params(m, mu)
        solve(ctx, b2, planned=False)
        ctx.set_params(m[::-1].copy(), mu)
        same(solve(ctx, b2, planned=True), ref3)
        ctx.set_cg(1e-30, 31)                           # max_iter inside the heads
        short = solve(ctx, b2, planned=True)
        assert np.all(short[1].status == tb.CG_MAXITER) and np.all(short[1].iters == 30)
        same(solve(ctx, b2, planned=False), short)
        ctx.set_cg(1e-30, 100000)
        ctx.set_params(m, mu)
        solve(ctx, b2, planned=False)
        x, info2 = solve(ctx, b2, planned=True)
        if nt * nx <= 128 * 128:
            c = C - 1
            xo, st, it, rr = oracle.fmdm_invert_cg(b2[c], A[c], float(m[c]), mu, tb.MODE_ADJOINT)
            assert st == info2.status[c] and abs(it - int(info2.iters[c])) <= 1
            assert_close(x[c], xo, CG_SOL_TOL, "planned cluster launch vs oracle")
