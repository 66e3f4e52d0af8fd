"""Thread-block-cluster CG (tb_cluster.cu: one cluster of 4096-site t-slabs per chain, halos and reductions over
distributed shared memory) against the CPU oracle and against the streaming solver, through the C-ABI.
Tolerances as everywhere: iteration count +-1, solution 1e-12 l2/max relative."""
import numpy as np
import pytest

from tests.util import CG_SOL_TOL, assert_close, random_gauge, random_vector

pytestmark = pytest.mark.gpu

tb = pytest.importorskip("thirring2d_b200")

CASES = [
    # nt, nx, nchains, mode, m, mu
    (128, 128, 3, tb.MODE_ADJOINT, 0.3, 0.0),     # 4 CTAs of 32 rows
    (64, 128, 2, tb.MODE_ADJOINT, 0.5, 0.1),      # 2 CTAs of 32 rows
    (128, 64, 2, tb.MODE_REF_COMPAT, 100.0, 0.1), # 2 CTAs of 64 rows, the shipped parameter regime (M~ = M)
    (256, 256, 2, tb.MODE_ADJOINT, 0.5, 0.0),     # 16 CTAs of 16 rows (non-portable cluster size)
    (128, 256, 1, tb.MODE_ADJOINT, 0.4, -0.2),    # 8 CTAs of 16 rows
    (256, 128, 1, tb.MODE_ADJOINT, 0.4, 0.0),     # 8 CTAs of 32 rows
    (256, 64, 2, tb.MODE_ADJOINT, 0.6, 0.05),     # 4 CTAs of 64 rows
    (64, 256, 2, tb.MODE_ADJOINT, 0.7, 0.1),      # 4 CTAs of 16 rows
]


@pytest.mark.parametrize("nt,nx,nchains,mode,m,mu", CASES)
def test_cluster_cg_matches_oracle_and_streaming(oracle, nt, nx, nchains, mode, m, mu):
    rng = np.random.default_rng(nt * 3 + nx + nchains)
    A = random_gauge(rng, nchains, nt, nx)
    xi = random_vector(rng, nchains, nt, nx)
    with tb.Context(nt, nx, nchains, mode, m=m, mu=mu) as ctx:
        ctx.set_gauge(A)
        b = ctx.fm_conjugate_mul(xi)
        ctx.set_tuning(solver=2)   # on-chip: must be the cluster kernel for these shapes, or fail loudly
        kind, in_flight = ctx.solver_info()
        assert kind == 2 and in_flight >= 1
        x, info = ctx.fmdm_invert_cg(b)
        x2, info2 = ctx.fmdm_invert_cg(b)
        ctx.set_tuning(solver=1)
        xs, infos = ctx.fmdm_invert_cg(b)
    assert np.array_equal(x, x2) and np.array_equal(info.iters, info2.iters)   # run-to-run deterministic
    for c in range(nchains):
        xo, st, it, rr = oracle.fmdm_invert_cg(b[c], A[c], m, mu, mode)
        assert info.status[c] == st == tb.CG_CONVERGED
        assert abs(int(info.iters[c]) - it) <= 1, (c, info.iters[c], it)
        assert abs(int(info.iters[c]) - int(infos.iters[c])) <= 1
        assert info.rr[c] < 1e-30
        assert_close(x[c], xo, CG_SOL_TOL, f"cluster vs oracle, chain {c}")
        assert_close(x[c], xs[c], CG_SOL_TOL, f"cluster vs streaming, chain {c}")


def test_cluster_more_chains_than_resident_clusters_and_ragged_masses():
    """80 chains of 128^2 (more than the ~33 clusters a B200 holds at once) with per-chain masses: later waves are
    scheduled as clusters retire, every chain converges at its own iteration and matches the streaming solver."""
    rng = np.random.default_rng(12)
    nt = nx = 128
    n = 80
    masses = rng.choice([0.2, 0.5, 1.0, 3.0], size=n)
    mus = rng.uniform(-0.1, 0.1, size=n)
    A = random_gauge(rng, n, nt, nx)
    xi = random_vector(rng, n, nt, nx)
    with tb.Context(nt, nx, n, tb.MODE_ADJOINT) as ctx:
        ctx.set_params(masses, mus)
        ctx.set_gauge(A)
        b = ctx.fm_conjugate_mul(xi)
        ctx.set_tuning(solver=2)
        x, info = ctx.fmdm_invert_cg(b)
        r = ctx.fmdm_mul(x) - b
        ctx.set_tuning(solver=1)
        xs, infos = ctx.fmdm_invert_cg(b)
    assert np.all(info.status == tb.CG_CONVERGED)
    assert np.all(np.abs(info.iters.astype(int) - infos.iters.astype(int)) <= 1)
    assert len(set(info.iters.tolist())) > 3
    assert_close(x, xs, CG_SOL_TOL, "cluster vs streaming")
    rel = np.linalg.norm(r.reshape(n, -1), axis=1) / np.linalg.norm(b.reshape(n, -1), axis=1)
    assert rel.max() < 1e-12, rel.max()


def test_cluster_zero_source_divergence_and_max_iter(oracle):
    rng = np.random.default_rng(9)
    nt = nx = 128
    A = random_gauge(rng, 3, nt, nx)
    b = random_vector(rng, 3, nt, nx)
    b[1] = 0.0  # hmc.c:359-361
    with tb.Context(nt, nx, 3, tb.MODE_ADJOINT, m=0.5) as ctx:
        ctx.set_gauge(A)
        ctx.set_tuning(solver=2)
        x, info = ctx.fmdm_invert_cg(b)
        assert info.status.tolist() == [tb.CG_CONVERGED, tb.CG_ZERO_SOURCE, tb.CG_CONVERGED]
        assert info.iters[1] == 0 and np.all(x[1] == 0)
        ctx.set_cg(1e-30, 20)
        x, info = ctx.fmdm_invert_cg(b)
        assert info.status.tolist() == [tb.CG_MAXITER, tb.CG_ZERO_SOURCE, tb.CG_MAXITER]
        assert info.iters.tolist() == [19, 0, 19]  # k = 1 .. max_iter-1, hmc.c:364
    # REF_COMPAT at light mass: M.M is not positive definite -> the reference bails (hmc.c:383-388)
    b[1] = b[0]
    with tb.Context(nt, nx, 3, tb.MODE_REF_COMPAT, m=0.1) as ctx:
        ctx.set_gauge(A)
        ctx.set_tuning(solver=2)
        x, info = ctx.fmdm_invert_cg(b)
        xo, st, it, rr = oracle.fmdm_invert_cg(b[0], A[0], 0.1, 0.0, tb.MODE_REF_COMPAT)
        assert st == tb.CG_DIVERGED
        assert info.status.tolist() == [tb.CG_DIVERGED] * 3
        assert abs(int(info.iters[0]) - it) <= 2


def test_cluster_device_resident_invert_and_condensate():
    """fm_invert_cg on 128^2 through the cluster solver: M (M^-1 v) = v, and the free-field condensate."""
    rng = np.random.default_rng(4)
    nt = nx = 128
    n = 6
    A = random_gauge(rng, n, nt, nx)
    v = random_vector(rng, n, nt, nx)
    with tb.Context(nt, nx, n, tb.MODE_ADJOINT, m=0.4, mu=0.1) as ctx:
        ctx.set_gauge(A)
        x, info = ctx.fm_invert_cg(v)
        assert np.all(info.status == tb.CG_CONVERGED)
        assert_close(ctx.fm_mul(x), v, 1e-11, "M M^-1 v")
    m = 0.5
    with tb.Context(nt, nx, n, tb.MODE_ADJOINT, m=m, mu=0.0) as ctx:
        ctx.set_gauge(np.zeros((n, nt, nx, 2)))
        cond, iters = ctx.hmc_condensate(nsrc=4, seed=3)
    k = (2 * np.arange(nt) + 1) * np.pi / nt
    s = np.sin(k) ** 2
    h = m / (m * m + s[:, None] + s[None, :])   # eigenvalues of the Hermitian part of M^-1 on the free field
    want = float(h.mean())
    # variance of the estimator for circular Gaussian sources with E|eta|^2 = 2: tr(H^2) / (V^2 nsrc) per chain (the
    # spread of six chains is too poor an estimate of it to gate on)
    err = float(np.sqrt((h * h).mean() / (nt * nx * 4) / n))
    assert abs(cond.mean() - want) < 4 * err, (cond.mean(), want, err)
    assert 0.2 * err * np.sqrt(n) < cond.std(ddof=1) < 3 * err * np.sqrt(n)   # chains draw different sources


@pytest.mark.parametrize("nt,nx", [(128, 64), (128, 128)])
def test_cluster_family_b_masked_operator(oracle, nt, nx):
    """Family B (vec_ops.c: real operator with an occupation mask) on the cluster kernel: cg_MdM and cg_propagator
    against the oracle, on-chip and streaming."""
    rng = np.random.default_rng(nt + nx)
    nsrc, m, mu = 3, 0.3, 0.1
    field = (rng.random((nsrc, nt, nx)) < 0.1).astype(np.int32)
    psi = rng.normal(size=(nsrc, nt, nx))
    with tb.Context(nt, nx, nsrc, tb.MODE_ADJOINT, m=m, mu=mu) as ctx:
        ctx.set_occupancy(field)
        ctx.set_tuning(solver=2)
        assert ctx.solver_info()[0] == 2
        x, info = ctx.cg_MdM(psi)
        xp, infop = ctx.cg_propagator(psi)
        ctx.set_tuning(solver=1)
        xs, infos = ctx.cg_MdM(psi)
    for c in range(nsrc):
        xo, st, it, rr = oracle.cg_MdM(psi[c], field[c], m, mu)
        assert info.status[c] == st == tb.CG_CONVERGED and abs(int(info.iters[c]) - it) <= 1
        assert_close(x[c], xo, CG_SOL_TOL, "cg_MdM")
        assert_close(x[c], xs[c], CG_SOL_TOL, "on-chip vs streaming")
        xo, st, it, rr = oracle.cg_MdM(psi[c], field[c], m, mu, propagator=True)
        assert abs(int(infop.iters[c]) - it) <= 1
        assert_close(xp[c], xo, CG_SOL_TOL, "cg_propagator")
        occ_sites = field[c] != 0
        assert np.allclose(x[c][occ_sites], psi[c][occ_sites], rtol=1e-12, atol=1e-13)


def test_cluster_hmc_trajectory_and_measurements():
    """The device-resident trajectory (hmc.c:671-746) on a 128^2 lattice: its 11*(nsteps/10)+1 solves run on the
    cluster kernel; observables are finite, exactly the accepted chains move, and a reversed-sign check of dS:
    the same trajectory replayed with the streaming solver gives the same observables to solver accuracy."""
    nt = nx = 128
    n = 5
    with tb.Context(nt, nx, n, tb.MODE_ADJOINT, m=0.5, mu=0.0) as ctx:
        ctx.hmc_set_coupling(0.3)
        ctx.hmc_heatbath(30, seed=21)
        assert ctx.solver_info()[0] == 2
        A0 = ctx.get_gauge()
        obs, acc, iters = ctx.hmc_trajectory(nsteps=10, traj_length=0.5, seed=5, traj_index=1)
        A1 = ctx.get_gauge()
        mag, ph = ctx.hmc_measure(nsrc=2, seed=5, meas_index=1)
        # replay from the same start with the streaming kernels
        ctx.set_gauge(A0)
        ctx.set_tuning(solver=1)
        obs_s, acc_s, iters_s = ctx.hmc_trajectory(nsteps=10, traj_length=0.5, seed=5, traj_index=1)
    assert np.all(np.isfinite(obs)) and iters > 0
    changed = np.abs(A1 - A0).reshape(n, -1).max(axis=1) > 0
    assert np.array_equal(changed, acc.astype(bool))
    assert np.all(np.isfinite(mag)) and np.all(np.isfinite(ph))
    assert np.array_equal(acc, acc_s)
    assert np.allclose(obs[:, :8], obs_s[:, :8], rtol=1e-9, atol=0)          # actions
    assert np.allclose(obs[:, 8], obs_s[:, 8], rtol=0, atol=1e-6 * np.abs(obs[:, :8]).max())   # dS (a difference)
