"""Golden record of BASELINE config 3 (256 x 256, ADJOINT, m = 0.01, g = 1: ~3 100 CG iterations) from the CPU oracle,
which tests/test_oracle_pinned.py pins bit for bit to the compiled reference (libhmcref_256x256_adjoint.so).  A reference
solve at this size is ~50 s of CPU, too long for the GPU box's test run, so the outcome is committed: iteration counts
(reference order and tree-summed dot products), final residual, ||x||^2 and a sample of the solution.

    python tests/golden/make_golden_256.py            (about 3 minutes)
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.pyoracle import MODE_ADJOINT, Oracle, RefLib, ref_available  # noqa: E402

NT = NX = 256
M, MU, G, SEED = 0.01, 0.0, 1.0, 2560001


def inputs():
    """Quenched-equilibrium links at coupling g (P(A) ~ exp((Nf/g) cos A), Nf = 2) and a Gaussian source."""
    rng = np.random.default_rng(SEED)
    A = rng.vonmises(0.0, 2.0 / G, size=(NT, NX, 2))
    xi = rng.normal(size=(NT, NX)) + 1j * rng.normal(size=(NT, NX))
    return A, xi


def main():
    orc = Oracle()
    A, xi = inputs()
    b = orc.fm_conjugate_mul(xi, A, M, MU, MODE_ADJOINT)
    x, st, it, rr = orc.fmdm_invert_cg(b, A, M, MU, MODE_ADJOINT)
    xt, stt, it_tree, rrt = orc.fmdm_invert_cg(b, A, M, MU, MODE_ADJOINT, treesum=True)
    same_as_reference = None
    if ref_available(NT, NX, "adjoint"):   # the compiled reference itself, when this container has it
        ref = RefLib(NT, NX, "adjoint", m=M, g=G, mu=MU)
        xr = ref.fmdm_invert_cg(b, ref.gauge(A))
        same_as_reference = bool(np.array_equal(xr, x))
    idx = np.arange(0, NT * NX, 997)
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "cg_256x256_m0.01_g1.npz")
    np.savez(out, nt=NT, nx=NX, m=M, mu=MU, g=G, seed=SEED, status=st, iters=it, iters_treesum=it_tree, rr=rr,
             x_norm2=float(np.vdot(x, x).real), sample_index=idx, x_sample=x.ravel()[idx],
             A_sum=float(A.sum()), b_norm2=float(np.vdot(b, b).real),
             tree_vs_sequential_rel_l2=float(np.linalg.norm(xt - x) / np.linalg.norm(x)),
             bitwise_equal_to_compiled_reference=-1 if same_as_reference is None else int(same_as_reference))
    print("iterations", it, "tree-summed", it_tree, "rr", rr, "bitwise equal to the compiled reference:", same_as_reference)


if __name__ == "__main__":
    main()
