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


def libm_cos_sin(A):
    """(cos A, sin A) evaluated by the C library's cos / sin (what hmc.c:140-174 calls), not by numpy's own vector
    routines: the last bit matters to the bit-parity tests of the strict solver."""
    import ctypes
    import ctypes.util

    libm = ctypes.CDLL(ctypes.util.find_library("m") or "libm.so.6")
    for f in (libm.cos, libm.sin):
        f.restype = ctypes.c_double
        f.argtypes = [ctypes.c_double]
    flat = np.ascontiguousarray(A, dtype=np.float64).ravel()
    out = np.empty((flat.size, 2))
    for i, a in enumerate(flat.tolist()):
        out[i, 0] = libm.cos(a)
        out[i, 1] = libm.sin(a)
    return out.reshape(np.shape(A) + (2,))


def iteration_band(oracle, b, A, m, mu, mode, it_ref):
    """Iterations by which the reference's OWN recursion moves when nothing but the summation order of its two dot
    products changes (sequential -> pairwise tree, oracle treesum variant).  A parallel solver necessarily sums as a
    tree; the +-1 of the north star is asserted on top of this measured, input-specific shift -- and the strict solver
    (reference summation order) must land on the reference's count itself (tests/test_gpu_strict.py)."""
    _, st, it_tree, _ = oracle.fmdm_invert_cg(b, A, m, mu, mode, treesum=True)
    return abs(it_tree - it_ref), it_tree
