"""BASELINE config 3 itself against the reference: 256 x 256, ADJOINT, m = 0.01, g = 1 (about 3 100 CG iterations) on
the cluster solver (plain and planned launch), the streaming solver and the strict solver, compared with the committed
outcome of the oracle (tests/golden/cg_256x256_m0.01_g1.npz from tests/golden/make_golden_256.py; the oracle is pinned
bit for bit to libhmcref_256x256_adjoint.so).  hmc.c:341-404."""
import os

import numpy as np
import pytest

from tests.golden.make_golden_256 import M, MU, NT, NX, inputs
from tests.util import libm_cos_sin

pytestmark = [pytest.mark.gpu, pytest.mark.slow]

tb = pytest.importorskip("thirring2d_b200")
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "cg_256x256_m0.01_g1.npz")


def test_config3_light_mass_solve_matches_the_reference(oracle, capsys):
    gold = np.load(GOLDEN)
    A, xi = inputs()
    b = oracle.fm_conjugate_mul(xi, A, M, MU, tb.MODE_ADJOINT)   # one CPU apply; the solve itself is the golden record
    assert abs(A.sum() - float(gold["A_sum"])) <= 1e-9 * abs(float(gold["A_sum"])) and \
        abs(np.vdot(b, b).real - float(gold["b_norm2"])) <= 1e-12 * float(gold["b_norm2"]), \
        "the seeded inputs differ from the ones the golden record was made with (numpy version?): regenerate it"
    it_ref, it_tree = int(gold["iters"]), int(gold["iters_treesum"])
    band = abs(it_tree - it_ref) + 1   # tests/util.py:iteration_band
    idx, xs_ref = gold["sample_index"], gold["x_sample"]
    n = 8   # 8 chains per GPU: 7 co-resident 16-CTA clusters, so the second solve is a planned launch
    An = np.broadcast_to(A, (n,) + A.shape)
    bn = np.broadcast_to(b, (n,) + b.shape)
    report = {}
    with tb.Context(NT, NX, n, tb.MODE_ADJOINT, m=M, mu=MU) as ctx:
        ctx.set_gauge(An)
        for name, solver in (("cluster", 2), ("cluster, planned launch", 2), ("streaming", 1)):
            ctx.set_tuning(solver=solver)
            x, info = ctx.fmdm_invert_cg(bn)
            assert np.all(info.status == tb.CG_CONVERGED)
            assert np.all(info.iters == info.iters[0]) and all(np.array_equal(x[c], x[0]) for c in range(1, n))
            d = int(info.iters[0]) - it_ref
            report[name] = d
            assert abs(d) <= band, (name, int(info.iters[0]), it_ref, it_tree)
            assert np.linalg.norm(x[0].ravel()[idx] - xs_ref) <= 1e-12 * np.linalg.norm(xs_ref)
            assert abs(np.vdot(x[0], x[0]).real - float(gold["x_norm2"])) <= 1e-12 * float(gold["x_norm2"])
    # the strict solver (reference evaluation order, host cos / sin): the reference's count and bits
    trig = libm_cos_sin(A)
    with tb.Context(NT, NX, 1, tb.MODE_ADJOINT, m=M, mu=MU) as ctx:
        ctx.set_links_trig(trig[..., 0, :], trig[..., 1, :])
        ctx.set_tuning(solver=5)
        x, info = ctx.fmdm_invert_cg(b)
    report["strict"] = int(info.iters[0]) - it_ref
    assert int(info.iters[0]) == it_ref and info.rr[0] == float(gold["rr"])
    assert np.array_equal(x.ravel()[idx], xs_ref)
    with capsys.disabled():
        print(f"\nconfig 3 (256x256, m=0.01, g=1): reference {it_ref} iterations, reference with tree-summed dot "
              f"products {it_tree}; GPU minus reference: {report}")
