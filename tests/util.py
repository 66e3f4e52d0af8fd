"""Shared helpers for the parity tests: seeded inputs and the tolerance rule of SURVEY Appendix C."""
import numpy as np

APPLY_TOL = 1e-13  # north_star: Dirac apply within 1e-13 relative (FP64)
CG_SOL_TOL = 1e-12  # norm-wise solution parity for m >= 0.01


def random_gauge(rng, nchains, nt, nx):
    """A_mu(t,x) i.i.d. uniform in [-pi, pi): the heat-bath proposal distribution (hmc.c:85)."""
    return rng.uniform(-np.pi, np.pi, size=(nchains, nt, nx, 2))


def smooth_gauge(rng, nchains, nt, nx, width):
    return rng.normal(0.0, width, size=(nchains, nt, nx, 2))


def random_vector(rng, nchains, nt, nx):
    """Complex N(0,1)+iN(0,1) per site, the distribution of stochastic_vector (hmc.c:439-447)."""
    return rng.normal(size=(nchains, nt, nx)) + 1j * rng.normal(size=(nchains, nt, nx))


def assert_close(got, ref, tol, what=""):
    """||d||_2/||ref||_2 <= tol and max|d| <= tol*max|ref| (element-wise relative error is not gated)."""
    got = np.asarray(got)
    ref = np.asarray(ref)
    d = got - ref
    n2 = np.linalg.norm(d.ravel()) / max(np.linalg.norm(ref.ravel()), 1e-300)
    mx = np.abs(d).max() / max(np.abs(ref).max(), 1e-300)
    assert n2 <= tol and mx <= tol, f"{what}: rel l2 {n2:.3e}, rel max {mx:.3e} > {tol:g}"
    return n2, mx
