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


# ---- tiles of 16 chains moved as 2-D boxes through tensor maps (batches of more than 16 chains) -------------------------
@pytest.fixture
def stage_tiled(monkeypatch):
    monkeypatch.setenv("TB_PIPE_TEST", "1")
    monkeypatch.setenv("TB_PIPE_TILED", "1")   # read when the context is created
    monkeypatch.delenv("TB_NO_PIPE", raising=False)
    monkeypatch.delenv("TB_PIPE_XPAY", raising=False)


TILED_CASES = [
    (32, 32, 64, 0.3, 0.0, (0, 17, 63)),     # 4 chain tiles x 2 site tiles; the marching kernels count in tiles of 32 chains
    (64, 64, 32, 0.2, 0.1, (0, 31)),         # 2 chain tiles, mu != 0
    (16, 48, 48, 0.4, 0.0, (0, 47)),         # 3 chain tiles x 3 site tiles, neither a power of two
    (24, 32, 128, 0.5, 0.05, (5, 127)),      # 8 chain tiles, 24 rows per block
    (32, 32, 20, 0.3, 0.0, (0, 15, 16, 19)),  # a ragged last tile: 4 of its 16 chains exist (20 sources on one field)
    (16, 48, 40, 0.4, 0.1, (0, 31, 32, 39)),  # 2 full tiles + 8 chains
]


@pytest.mark.parametrize("nt,nx,C,m,mu,check", TILED_CASES)
def test_tiled_staged_kernels_match_the_oracle_and_the_marching_kernels(stage_tiled, oracle, nt, nx, C, m, mu, check):
    rng = np.random.default_rng(nt * 11 + nx + C)
    A = random_gauge(rng, C, nt, nx)
    xi = random_vector(rng, C, nt, nx)
    xi[C // 2] = 0.0                                    # a zero source: its chain is masked from the start
    masses = np.full(C, m)
    masses[3] = 4 * m                                   # a chain that finishes long before the others
    masses[16:32] = 3 * m                               # a whole 16-chain tile (or what exists of it) that finishes early
    with tb.Context(nt, nx, C, tb.MODE_ADJOINT, m=masses, mu=mu) as ctx:
        ctx.set_tuning(solver=1)
        assert ctx.streaming_info()[:3] == (2, 16, 16), ctx.streaming_info()
        ctx.set_gauge(A)
        # T1 for the tiled plain apply
        v = random_vector(rng, C, nt, nx)
        got_m, got_d = ctx.fm_mul(v), ctx.fm_dagger_mul(v)
        for c in check:
            assert_close(got_m[c], oracle.fm_mul(v[c], A[c], float(masses[c]), mu), 1e-13, f"tiled M, chain {c}")
            assert_close(got_d[c], oracle.fm_dagger_mul(v[c], A[c], float(masses[c]), mu), 1e-13, f"tiled M^dagger, chain {c}")
        b = ctx.fm_conjugate_mul(xi)
        xs, is_, ls = solve(ctx, b, staged=True)
        xm, im, lm = solve(ctx, b, staged=False)
        assert np.array_equal(is_.status, im.status)
        assert np.all(np.abs(is_.iters.astype(int) - im.iters.astype(int)) <= 1), (is_.iters, im.iters)
        for c in range(C):
            if is_.status[c] == tb.CG_CONVERGED:
                assert_close(xs[c], xm[c], CG_SOL_TOL, f"tiled vs marching, chain {c}")
            else:
                assert is_.status[c] == tb.CG_ZERO_SOURCE and not xs[c].any()
        for c in check:
            xo, st, it, rr = oracle.fmdm_invert_cg(b[c], A[c], float(masses[c]), mu, tb.MODE_ADJOINT)
            assert st == is_.status[c] and abs(it - int(is_.iters[c])) <= 1, (c, it, is_.iters[c])
            assert_close(xs[c], xo, CG_SOL_TOL, f"tiled kernels vs oracle, chain {c}")
        xs2, is2, _ = solve(ctx, b, staged=True)
        assert np.array_equal(xs, xs2) and np.array_equal(is_.iters, is2.iters)


def test_tiled_kernels_are_the_default_for_a_batch_of_more_than_16_chains(monkeypatch):
    """256^2 x 32 chains, no switches set: the streaming solver stages 16-chain tiles by itself and agrees with the
    marching kernels."""
    for k in ("TB_PIPE_TEST", "TB_PIPE_TILED", "TB_NO_PIPE", "TB_PIPE_XPAY"):
        monkeypatch.delenv(k, raising=False)
    nt = nx = 256
    C = 32
    rng = np.random.default_rng(6)
    A = smooth_gauge(rng, C, nt, nx, 0.5)
    xi = random_vector(rng, C, nt, nx)
    with tb.Context(nt, nx, C, tb.MODE_ADJOINT, m=0.3, mu=0.0) as ctx:
        ctx.set_tuning(solver=1)
        assert ctx.streaming_info()[:3] == (2, 16, 16), ctx.streaming_info()
        ctx.set_gauge(A)
        b = ctx.fm_conjugate_mul(xi)
        xs, is_, _ = solve(ctx, b, staged=True)
        xm, im, _ = solve(ctx, b, staged=False)
        assert np.all(is_.status == tb.CG_CONVERGED) and np.all(np.abs(is_.iters.astype(int) - im.iters.astype(int)) <= 1)
        for c in (0, 15, 16, 31):
            assert_close(xs[c], xm[c], CG_SOL_TOL, f"chain {c}")
    with tb.Context(512, 512, 16, tb.MODE_ADJOINT, m=0.3, mu=0.0) as ctx:   # 16 chains: one tile holds the whole batch
        ctx.set_tuning(solver=1)
        assert ctx.streaming_info()[:3] == (1, 16, 16), ctx.streaming_info()
    with tb.Context(nt, nx, 24, tb.MODE_ADJOINT, m=0.3, mu=0.0) as ctx:     # not a multiple of 16: a ragged last tile
        ctx.set_tuning(solver=1)
        assert ctx.streaming_info()[:3] == (2, 16, 16), ctx.streaming_info()
    with tb.Context(nt, nx, 12, tb.MODE_ADJOINT, m=0.3, mu=0.0) as ctx:     # fewer than 16, not a power of two: marching
        ctx.set_tuning(solver=1)
        assert ctx.streaming_info()[0] == 0, ctx.streaming_info()


# ---- the direction update folded into the first staged pass: two launches per iteration ---------------------------------
@pytest.mark.parametrize("nt,nx,C,m,mu,tiled", [(64, 128, 8, 0.2, 0.0, False), (128, 256, 1, 0.3, 0.05, False),
                                                (24, 64, 4, 0.4, 0.0, False), (16, 32, 128, 0.5, 0.0, False),
                                                (32, 32, 64, 0.3, 0.0, True), (16, 48, 48, 0.4, 0.1, True)])
def test_two_launch_iteration_is_bitwise_the_three_launch_one(monkeypatch, nt, nx, C, m, mu, tiled):
    """p = r + beta p formed inside the first pass is the same fma the xpay kernel does and the sums run over the same
    blocks: solution, iteration counts and residuals are bit for bit those of the three-launch staged iteration."""
    monkeypatch.setenv("TB_PIPE_TEST", "1")
    monkeypatch.setenv("TB_PIPE_TILED", "1" if tiled else "0")
    monkeypatch.delenv("TB_NO_PIPE", raising=False)
    rng = np.random.default_rng(nt + nx + C)
    A = random_gauge(rng, C, nt, nx)
    xi = random_vector(rng, C, nt, nx)
    masses = np.full(C, m)
    if C >= 8:
        masses[3] = 4 * m
        xi[C // 2] = 0.0
    with tb.Context(nt, nx, C, tb.MODE_ADJOINT, m=masses, mu=mu) as ctx:
        ctx.set_tuning(solver=1)
        assert ctx.streaming_info()[0] == (2 if tiled else 1)
        ctx.set_gauge(A)
        b = ctx.fm_conjugate_mul(xi)
        res = {}
        for xpay in ("0", "1"):
            monkeypatch.setenv("TB_PIPE_XPAY", xpay)
            n0 = ctx.launch_count
            x, info = ctx.fmdm_invert_cg(b)
            res[xpay] = (x, info, ctx.launch_count - n0)
        (x3, i3, l3), (x2, i2, l2) = res["0"], res["1"]
        assert np.array_equal(i3.iters, i2.iters) and np.array_equal(i3.status, i2.status)
        assert np.array_equal(i3.rr, i2.rr)
        assert np.array_equal(x3, x2)
        assert l2 < l3 and (l3 - l2) * 4 >= l3   # two launches per iteration instead of three
        # and again, starting from the other parity of the two direction buffers' history
        x2b, i2b = ctx.fmdm_invert_cg(b)
        assert np.array_equal(x2b, x2) and np.array_equal(i2b.iters, i2.iters)


# ---- one gauge field shared by every chain of the batch (multi-RHS): compact links in the staged kernels ---------------
@pytest.mark.parametrize("nt,nx,C,mu,tiled,xpay", [(64, 128, 8, 0.0, False, "0"), (64, 64, 16, 0.1, False, "1"),
                                                   (32, 32, 64, 0.0, True, "0"), (16, 48, 40, 0.1, True, "1"),
                                                   (32, 32, 20, 0.0, True, "0")])
def test_shared_gauge_field_is_bitwise_the_replicated_one(monkeypatch, oracle, nt, nx, C, mu, tiled, xpay):
    """tb_set_gauge_shared: the staged kernels read one link per site instead of one per site and chain; the arithmetic
    per site is untouched, so applies are bit for bit those of the same field uploaded once per chain with tb_set_gauge,
    and so is the solve on the marching kernels and on the on-chip solver (which use the per-chain copies the call also
    leaves behind); the staged solve agrees to rounding (its blocks are shorter with a shared field)."""
    monkeypatch.setenv("TB_PIPE_TEST", "1")
    monkeypatch.setenv("TB_PIPE_TILED", "1" if tiled else "0")
    monkeypatch.setenv("TB_PIPE_XPAY", xpay)
    monkeypatch.delenv("TB_NO_PIPE", raising=False)
    rng = np.random.default_rng(3 * nt + nx + C)
    A1 = random_gauge(rng, 1, nt, nx)[0]
    A = np.ascontiguousarray(np.broadcast_to(A1, (C, nt, nx, 2)))
    xi = random_vector(rng, C, nt, nx)
    masses = np.full(C, 0.3)
    masses[3] = 1.2
    with tb.Context(nt, nx, C, tb.MODE_ADJOINT, m=masses, mu=mu) as ctx:
        ctx.set_tuning(solver=1)
        assert ctx.streaming_info()[0] == (2 if tiled else 1)
        res = {}
        for name in ("shared", "replicated", "shared again"):
            if name == "replicated":
                ctx.set_gauge(A)
            else:
                ctx.set_gauge_shared(A1)
            m_v, d_v = ctx.fm_mul(xi), ctx.fm_dagger_mul(xi)
            b = ctx.fm_conjugate_mul(xi)
            x, info = ctx.fmdm_invert_cg(b)
            os.environ["TB_NO_PIPE"] = "1"
            try:
                xm, im = ctx.fmdm_invert_cg(b)
            finally:
                os.environ.pop("TB_NO_PIPE", None)
            res[name] = (m_v, d_v, x, info, xm, im)
        ref = res["replicated"]
        for name in ("shared", "shared again"):
            got = res[name]
            for k in (0, 1, 4):          # the applies and the solve on the marching kernels: bit for bit
                assert np.array_equal(got[k], ref[k]), (name, k)
            assert np.array_equal(got[5].iters, ref[5].iters)
            # the staged solve runs shorter blocks with a shared field (tb_gauge_sharing): same arithmetic per site,
            # differently blocked sums
            assert np.array_equal(got[3].status, ref[3].status)
            assert np.all(np.abs(got[3].iters.astype(int) - ref[3].iters.astype(int)) <= 1)
            assert_close(got[2], ref[2], CG_SOL_TOL, f"{name} vs replicated, staged solve")
        assert np.array_equal(res["shared"][2], res["shared again"][2])
        c = C - 1
        b = ctx.fm_conjugate_mul(xi)
        xo, st, it, rr = oracle.fmdm_invert_cg(b[c], A1, float(masses[c]), mu, tb.MODE_ADJOINT)
        assert abs(it - int(res["shared"][3].iters[c])) <= 1
        assert_close(res["shared"][2][c], xo, CG_SOL_TOL, "shared gauge field vs oracle")


def test_shared_gauge_field_on_the_on_chip_solver(oracle):
    """64^2 x 20 sources on one field through the default (on-chip) solver: the per-chain link copies are right."""
    nt = nx = 64
    C = 20
    rng = np.random.default_rng(11)
    A1 = random_gauge(rng, 1, nt, nx)[0]
    xi = random_vector(rng, C, nt, nx)
    with tb.Context(nt, nx, C, tb.MODE_ADJOINT, m=0.2, mu=0.0) as ctx:
        ctx.set_gauge_shared(A1)
        b = ctx.fm_conjugate_mul(xi)
        x, info = ctx.fmdm_invert_cg(b)
        ctx.set_gauge(np.ascontiguousarray(np.broadcast_to(A1, (C, nt, nx, 2))))
        x2, info2 = ctx.fmdm_invert_cg(b)
        assert np.array_equal(x, x2) and np.array_equal(info.iters, info2.iters)
        xo, st, it, rr = oracle.fmdm_invert_cg(b[7], A1, 0.2, 0.0, tb.MODE_ADJOINT)
        assert abs(it - int(info.iters[7])) <= 1
        assert_close(x[7], xo, CG_SOL_TOL, "on-chip solver, shared gauge field")
