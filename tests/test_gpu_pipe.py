"""TMA-staged streaming kernels (tb_stream.cu: dslash_pipe_kernel): the two stencil passes of the fused CG iteration
with the rows of a block travelling through a ring of shared-memory stages (cp.async.bulk + mbarrier).  Same
arithmetic per site as the register-marching kernels; the per-chain sums are taken over differently shaped blocks, so
the comparison is the T4 one: iteration count within +-1 of the oracle, solution to 1e-12."""
import os

import numpy as np
import pytest

from tests.util import CG_SOL_TOL, assert_close, random_gauge, random_vector, smooth_gauge

pytestmark = pytest.mark.gpu

tb = pytest.importorskip("thirring2d_b200")


@pytest.fixture
def stage_small_lattices(monkeypatch):
    monkeypatch.setenv("TB_PIPE_TEST", "1")   # read when the context is created: every shape the kernels can handle
    monkeypatch.delenv("TB_NO_PIPE", raising=False)


def solve(ctx, b, staged):
    if staged:
        os.environ.pop("TB_NO_PIPE", None)
    else:
        os.environ["TB_NO_PIPE"] = "1"
    try:
        n0 = ctx.launch_count
        x, info = ctx.fmdm_invert_cg(b)
        return x, info, ctx.launch_count - n0
    finally:
        os.environ.pop("TB_NO_PIPE", None)


# nt, nx, chains, m, mu, chains checked against the oracle
CASES = [
    (32, 32, 64, 0.3, 0.0, (0, 63)),      # 64 chains x 4 sites per tile (two of the chain tiles tile_active counts in)
    (64, 64, 32, 0.2, 0.1, (0, 31)),      # 32 chains x 8 sites
    (64, 128, 8, 0.2, 0.0, (0, 7)),       # 8 chains x 32 sites per tile
    (128, 256, 1, 0.3, 0.05, (0,)),       # a single lattice: 256 sites per tile
    (16, 32, 128, 0.5, 0.0, (0, 127)),    # 128 chains x 2 sites, three stages in the fused pass, 16 rows
    (24, 64, 4, 0.4, 0.0, (1,)),          # rows per block = 24
]


@pytest.mark.parametrize("nt,nx,C,m,mu,check", CASES)
def test_staged_kernels_match_the_oracle_and_the_marching_kernels(stage_small_lattices, oracle, nt, nx, C, m, mu, check):
    rng = np.random.default_rng(nt * 7 + nx + C)
    A = random_gauge(rng, C, nt, nx)
    xi = random_vector(rng, C, nt, nx)
    xi[C // 2] = 0.0 if C > 2 else xi[C // 2]          # a zero source: its chain is masked from the start
    masses = np.full(C, m)
    if C >= 8:
        masses[3] = 4 * m                               # a chain that finishes long before the others
    with tb.Context(nt, nx, C, tb.MODE_ADJOINT, m=masses, mu=mu) as ctx:
        ctx.set_tuning(solver=1)
        ctx.set_gauge(A)
        b = ctx.fm_conjugate_mul(xi)
        xs, is_, ls = solve(ctx, b, staged=True)
        xm, im, lm = solve(ctx, b, staged=False)
        assert np.array_equal(is_.status, im.status)
        assert np.all(np.abs(is_.iters.astype(int) - im.iters.astype(int)) <= 1), (is_.iters, im.iters)
        for c in range(C):
            if is_.status[c] == tb.CG_CONVERGED:
                assert_close(xs[c], xm[c], CG_SOL_TOL, f"staged vs marching, chain {c}")
            else:
                assert is_.status[c] == tb.CG_ZERO_SOURCE and not xs[c].any()
        for c in check:
            xo, st, it, rr = oracle.fmdm_invert_cg(b[c], A[c], float(masses[c]), mu, tb.MODE_ADJOINT)
            assert st == is_.status[c] and abs(it - int(is_.iters[c])) <= 1, (c, it, is_.iters[c])
            assert_close(xs[c], xo, CG_SOL_TOL, f"staged kernels vs oracle, chain {c}")
        # the staged solve repeated is bitwise itself (deterministic sums)
        xs2, is2, _ = solve(ctx, b, staged=True)
        assert np.array_equal(xs, xs2) and np.array_equal(is_.iters, is2.iters)


def test_staged_kernels_are_the_default_on_a_large_lattice():
    """512^2 x 16 chains: the streaming solver picks the staged kernels by itself (TB_NO_PIPE unset) and agrees with
    the marching kernels."""
    nt = nx = 512
    C = 16
    rng = np.random.default_rng(5)
    A = smooth_gauge(rng, C, nt, nx, 0.5)
    xi = random_vector(rng, C, nt, nx)
    with tb.Context(nt, nx, C, tb.MODE_ADJOINT, m=0.3, mu=0.0) as ctx:
        ctx.set_tuning(solver=1)
        ctx.set_gauge(A)
        b = ctx.fm_conjugate_mul(xi)
        xs, is_, _ = solve(ctx, b, staged=True)
        xm, im, _ = solve(ctx, b, staged=False)
        assert np.all(is_.status == tb.CG_CONVERGED) and np.all(np.abs(is_.iters.astype(int) - im.iters.astype(int)) <= 1)
        for c in (0, 7, 15):
            assert_close(xs[c], xm[c], CG_SOL_TOL, f"chain {c}")


@pytest.mark.parametrize("nt,nx,C,mu", [(32, 32, 64, 0.0), (64, 128, 8, 0.1), (128, 256, 1, 0.05), (16, 32, 128, 0.0),
                                        (24, 64, 4, 0.2)])
def test_staged_apply_matches_oracle(stage_small_lattices, oracle, nt, nx, C, mu):
    """T1 for the staged plain apply (CG = false): M and M^dagger against the oracle to 1e-13, every tile shape."""
    from tests.util import APPLY_TOL
    rng = np.random.default_rng(nt + 3 * nx + C)
    A = random_gauge(rng, C, nt, nx)
    v = random_vector(rng, C, nt, nx)
    m = rng.uniform(0.05, 1.0, size=C)
    with tb.Context(nt, nx, C, tb.MODE_ADJOINT, m=m, mu=mu) as ctx:
        ctx.set_gauge(A)
        got_m, got_d = ctx.fm_mul(v), ctx.fm_dagger_mul(v)
        os.environ["TB_NO_PIPE"] = "1"
        try:
            ref_m, ref_d = ctx.fm_mul(v), ctx.fm_dagger_mul(v)
        finally:
            os.environ.pop("TB_NO_PIPE", None)
        for c in sorted({0, C // 2, C - 1}):
            assert_close(got_m[c], oracle.fm_mul(v[c], A[c], float(m[c]), mu), 1e-13, f"staged M, chain {c}")
            assert_close(got_d[c], oracle.fm_dagger_mul(v[c], A[c], float(m[c]), mu), 1e-13, f"staged M^dagger, chain {c}")
        assert_close(got_m, ref_m, APPLY_TOL, "staged vs marching M")
        assert_close(got_d, ref_d, APPLY_TOL, "staged vs marching M^dagger")
