"""The strict solver (tb_set_tuning solver = 5): fmdm_invert_cg in the reference's own floating-point evaluation order
-- no FMA contraction, lattice sums accumulated sequentially in (t, x) order (hmc.c:354-379) -- with the links' cos / sin
from the host's libm.  It must reproduce the reference's recursion bit for bit: the SAME iteration count (not +-1) and
the same solution, at the light masses where the fast solvers' tree sums move the count by 1-3 (north star: "same
residual in the same iteration count").  The oracle it is compared with is pinned bitwise to the compiled reference
(tests/test_oracle_pinned.py)."""
import numpy as np
import pytest

from tests.util import iteration_band, libm_cos_sin, random_gauge, random_vector

pytestmark = pytest.mark.gpu

tb = pytest.importorskip("thirring2d_b200")

CASES = [
    # nt, nx, nchains, mode, m, mu
    (32, 32, 2, tb.MODE_ADJOINT, 0.01, 0.0),      # ~750 iterations; tree sums shift these by 2-24
    (64, 64, 2, tb.MODE_ADJOINT, 0.01, 0.0),      # ~2900 iterations
    (16, 32, 3, tb.MODE_ADJOINT, 0.05, 0.15),     # rectangle, chemical potential
    (32, 32, 1, tb.MODE_REF_COMPAT, 100.0, 0.1),  # the shipped parameter file regime (M~ = M)
    (64, 64, 1, tb.MODE_ADJOINT, 0.1, 0.0),       # the headline workload's mass
]


@pytest.mark.parametrize("nt,nx,nchains,mode,m,mu", CASES)
def test_strict_solver_reproduces_the_reference_recursion(oracle, nt, nx, nchains, mode, m, mu):
    rng = np.random.default_rng(4000 + nt + nx + nchains)
    A = random_gauge(rng, nchains, nt, nx)
    xi = random_vector(rng, nchains, nt, nx)
    b = np.stack([oracle.fm_conjugate_mul(xi[c], A[c], m, mu, mode) for c in range(nchains)])
    trig = libm_cos_sin(A)   # (nchains, nt, nx, dir, 2)
    with tb.Context(nt, nx, nchains, mode, m=m, mu=mu) as ctx:
        ctx.set_links_trig(trig[..., 0, :], trig[..., 1, :])
        ctx.set_tuning(solver=5)
        x, info = ctx.fmdm_invert_cg(b)
        ctx.set_tuning(solver=0)
        xf, infof = ctx.fmdm_invert_cg(b)   # the fast solver on the same links
    shifts = []
    for c in range(nchains):
        xo, st, it, rr = oracle.fmdm_invert_cg(b[c], A[c], m, mu, mode)
        assert info.status[c] == st == tb.CG_CONVERGED
        assert int(info.iters[c]) == it, ("strict solver", c, int(info.iters[c]), it)
        assert info.rr[c] == rr, ("final ||r||^2", info.rr[c], rr)
        assert np.array_equal(x[c], xo), f"chain {c}: max |dx| = {np.abs(x[c] - xo).max():.3e}"
        # the fast solver: the same solve to rounding, its count inside the tree-summation band of this input
        shift, it_tree = iteration_band(oracle, b[c], A[c], m, mu, mode, it)
        assert abs(int(infof.iters[c]) - it) <= shift + 1, (int(infof.iters[c]), it, it_tree)
        assert np.linalg.norm(xf[c] - xo) <= 1e-12 * np.linalg.norm(xo)
        shifts.append((it, it_tree, int(infof.iters[c])))
    print(f"{nt}x{nx} m={m}: (reference = strict, reference with tree sums, fast solver) = {shifts}")


def test_strict_solver_with_device_sincos_stays_within_one_iteration(oracle):
    """Without host cos / sin the links come from the device's sincos (<= 1-2 ulp from glibc's): still the reference's
    summation order, so at a few hundred iterations the count is the reference's +-1 and the solution agrees to 1e-13."""
    nt = nx = 32
    rng = np.random.default_rng(77)
    A = random_gauge(rng, 2, nt, nx)
    xi = random_vector(rng, 2, nt, nx)
    with tb.Context(nt, nx, 2, tb.MODE_ADJOINT, m=0.1, mu=0.05) as ctx:
        ctx.set_gauge(A)
        ctx.set_tuning(solver=5)
        b = ctx.fm_conjugate_mul(xi)
        x, info = ctx.fmdm_invert_cg(b)
    for c in range(2):
        xo, st, it, rr = oracle.fmdm_invert_cg(b[c], A[c], 0.1, 0.05, tb.MODE_ADJOINT)
        assert abs(int(info.iters[c]) - it) <= 1
        assert np.linalg.norm(x[c] - xo) <= 1e-13 * np.linalg.norm(xo)
