"""Pins the CPU oracle: (i) against the committed golden fixtures generated from the reference itself
(tests/golden/make_golden.py), (ii) bit-for-bit against the compiled reference when oracle/_ref is present.
CPU only."""
import glob
import os
import subprocess

import numpy as np
import pytest

from oracle import pyoracle
from oracle.pyoracle import CG_CONVERGED, CG_DIVERGED, MODE_ADJOINT, MODE_REF_COMPAT, RefLib, ref_available

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
FIXTURES = sorted(glob.glob(os.path.join(GOLD, "ref_*.npz")))


def test_fixtures_present():
    assert len(FIXTURES) >= 9


@pytest.mark.parametrize("path", FIXTURES, ids=[os.path.basename(p) for p in FIXTURES])
def test_oracle_reproduces_reference_fixture_bitwise(oracle, path):
    d = np.load(path)
    A, v, m, mu, mode = d["A"], d["v"], float(d["m"]), float(d["mu"]), int(d["mode"])
    assert np.array_equal(oracle.fm_mul(v, A, m, mu), d["Mv"])
    assert np.array_equal(oracle.fm_conjugate_mul(v, A, m, mu, mode), d["Mcv"])
    if mode == MODE_REF_COMPAT:
        assert np.array_equal(d["Mv"], d["Mcv"])  # the shipped "conjugate" is a copy of fm_mul
    if "x" in d:
        x, st, it, rr = oracle.fmdm_invert_cg(d["Mcv"], A, m, mu, mode)
        assert st == CG_CONVERGED and rr < 1e-30
        assert np.array_equal(x, d["x"])
    else:
        x, st, it, rr = oracle.fmdm_invert_cg(d["Mcv"], A, m, mu, mode)
        assert st == CG_DIVERGED  # the reference prints "Cannot invert fermion matrix" and exits here
    if "dense_row0" in d:
        D = oracle.fermion_matrix(A, m, mu)
        assert np.array_equal(D[0], d["dense_row0"]) and np.array_equal(D[:, 0], d["dense_col0"])
        # matrix-free apply against the reference's dense statement of the operator (hmc.c:269-310)
        assert np.abs(d["dense_Dv"] - d["Mv"].ravel()).max() <= 1e-15 * np.abs(d["Mv"]).max() * 8
        # true adjoint against conj-transpose of the dense matrix
        Mdv = oracle.fm_dagger_mul(v, A, m, mu)
        ref = D.conj().T @ v.ravel()
        assert np.abs(ref - Mdv.ravel()).max() <= 1e-15 * np.abs(ref).max() * 8


def test_oracle_adjoint_identity(oracle):
    """<a, M b> = <M^dagger a, b> (the corrected identity of SURVEY Appendix A.13)."""
    rng = np.random.default_rng(0)
    nt, nx = 12, 20
    A = rng.uniform(-np.pi, np.pi, size=(nt, nx, 2))
    a = rng.normal(size=(nt, nx)) + 1j * rng.normal(size=(nt, nx))
    b = rng.normal(size=(nt, nx)) + 1j * rng.normal(size=(nt, nx))
    lhs = np.vdot(a, oracle.fm_mul(b, A, 0.3, 0.2))
    rhs = np.vdot(oracle.fm_dagger_mul(a, A, 0.3, 0.2), b)
    assert abs(lhs - rhs) < 1e-12 * abs(lhs)


def test_oracle_free_field_condensate(oracle):
    """(1/V) Tr M^-1 at A = 0, mu = 0 equals the antiperiodic momentum sum."""
    L, m = 8, 0.5
    A = np.zeros((L, L, 2))
    D = oracle.fermion_matrix(A, m, 0.0)
    tr = np.trace(np.linalg.inv(D)).real / (L * L)
    k = (2 * np.arange(L) + 1) * np.pi / L
    s2 = np.sin(k) ** 2
    expect = np.mean(m / (m * m + s2[:, None] + s2[None, :]))
    assert abs(tr - expect) < 1e-12
    assert abs(tr - 0.4941176470588) < 1e-10


needs_ref = pytest.mark.skipif(not ref_available(16, 16, "compat"), reason="oracle/_ref not built")


@needs_ref
@pytest.mark.parametrize("nt,nx", [(8, 8), (16, 32), (32, 32)])
@pytest.mark.parametrize("flavour", ["compat", "adjoint"])
@pytest.mark.parametrize("m,mu", [(100.0, 0.1), (1.0, 0.3)])
def test_oracle_bitwise_vs_compiled_reference(oracle, nt, nx, flavour, m, mu):
    ref = RefLib(nt, nx, flavour, m=m, g=0.3, mu=mu, seed=12345 + nt)
    G = ref.gauge()
    ref.heatbath(G, 3)
    v = ref.stochastic_vector()
    assert np.array_equal(ref.fm_mul(v, G), oracle.fm_mul(v, G.A, m, mu))
    b = ref.fm_conjugate_mul(v, G)
    assert np.array_equal(b, oracle.fm_conjugate_mul(v, G.A, m, mu, ref.mode))
    assert np.array_equal(ref.fermion_matrix(G), oracle.fermion_matrix(G.A, m, mu))
    if flavour == "adjoint" or m == 100.0:
        x, st, it, rr = oracle.fmdm_invert_cg(b, G.A, m, mu, ref.mode)
        assert st == CG_CONVERGED
        assert np.array_equal(ref.fmdm_invert_cg(b, G), x)
        xi, st, it, rr = oracle.fm_invert_cg(v, G.A, m, mu, ref.mode)
        assert np.array_equal(ref.fm_invert_cg(v, G), xi)


@needs_ref
def test_reference_stdout_known_answer():
    """The reference build itself reproduces the stdout head recorded in SURVEY section 6."""
    params = "5\n1\n100\n0.3\n0.1\n4354365264\n"
    out = subprocess.run([os.path.join(pyoracle.REF_DIR, "ref_hmc"),
                          os.path.join(pyoracle.REF_DIR, "libhmcref_32x32_compat.so")],
                         input=params, capture_output=True, text=True, check=True).stdout
    with open(os.path.join(GOLD, "hmc_32x32_shipped_5traj.stdout")) as f:
        assert out == f.read()
    assert "Start HMC: Sg 1042.01, Smdm 2199.41, Smd 203252, Smom 2015.63" in out
    assert "HMC End, dS -2.75473, Sg 1115.92, Smdm 2198.94, Smd 203250, Sm 1941.25" in out
    assert "Phase -20.0647" in out


# ---- family B (vec_ops.c) --------------------------------------------------------------------------------
from oracle.pyoracle import RefLibB, ref_b_available  # noqa: E402

needs_ref_b = pytest.mark.skipif(not ref_b_available(16, 16), reason="oracle/_ref not built")


@needs_ref_b
@pytest.mark.parametrize("nt,nx", [(16, 16), (16, 32), (64, 64)])
@pytest.mark.parametrize("m,mu,occ", [(0.3, 0.1, 0.1), (0.05, 0.0, 0.0), (1.0, 0.3, 0.3)])
def test_family_b_oracle_bitwise_vs_compiled_vec_ops(oracle, nt, nx, m, mu, occ):
    rng = np.random.default_rng(nt * nx + int(100 * m))
    ref = RefLibB(nt, nx, m=m, mu=mu)
    field = (rng.random((nt, nx)) < occ).astype(np.int32)
    ref.set_field(field)
    psi = rng.normal(size=(nt, nx))
    assert np.array_equal(ref.call("fM", psi), oracle.fM(psi, field, m, mu))
    assert np.array_equal(ref.call("fM_transpose", psi), oracle.fM(psi, field, m, mu, transpose=True))
    x, st, it, rr = oracle.cg_MdM(psi, field, m, mu)
    assert st == CG_CONVERGED and np.array_equal(ref.call("cg_MdM", psi), x)
    xp, st, it, rr = oracle.cg_MdM(psi, field, m, mu, propagator=True)
    assert np.array_equal(ref.call("cg_propagator", psi), xp)
    # the reference's own sanity identities: <a, M b> = <M^T a, b>, M M^-1 b = b
    a = rng.normal(size=(nt, nx))
    assert abs(np.vdot(a, oracle.fM(psi, field, m, mu)) - np.vdot(oracle.fM(a, field, m, mu, transpose=True), psi)) < 1e-11
    assert np.abs(oracle.fM(xp, field, m, mu) - psi).max() < 1e-12


@pytest.mark.skipif(not ref_b_available(64, 64, 1), reason="oracle/_ref (SYMMETRIC build) not present")
@pytest.mark.parametrize("nt,nx", [(16, 32), (64, 64)])
@pytest.mark.parametrize("bc", [1, 2], ids=["SYMMETRIC", "OPENX"])
@pytest.mark.parametrize("m,mu,occ", [(0.3, 0.1, 0.1), (0.2, 0.0, 0.0)])
def test_family_b_boundary_variants_bitwise_vs_compiled_vec_ops(oracle, nt, nx, bc, m, mu, occ):
    """Thirring.h:27-29.  SYMMETRIC: vec_ops.c:175-249 from a build with the #define swapped; OPENX: the ANTISYMMETRIC
    object with the neighbour tables and the EMPTY phantom column fermionbag.c:713-717,761-765 sets up."""
    rng = np.random.default_rng(nt * nx + bc)
    ref = RefLibB(nt, nx, m=m, mu=mu, bc=bc)
    field = (rng.random((nt, nx)) < occ).astype(np.int32)
    ref.set_field(field)
    psi = rng.normal(size=(nt, nx))
    antisymmetric = oracle.fM(psi, field, m, mu)
    oracle.set_boundary(bc)
    try:
        assert np.array_equal(ref.call("fM", psi), oracle.fM(psi, field, m, mu))
        assert np.array_equal(ref.call("fM_transpose", psi), oracle.fM(psi, field, m, mu, transpose=True))
        assert not np.array_equal(oracle.fM(psi, field, m, mu), antisymmetric)   # the variant does change the operator
        xp, st, it, rr = oracle.cg_MdM(psi, field, m, mu, propagator=True)
        assert st == CG_CONVERGED and np.array_equal(ref.call("cg_propagator", psi), xp)
    finally:
        oracle.set_boundary(0)


def test_family_b_golden_fixture(oracle):
    d = np.load(os.path.join(GOLD, "refB_32x32_m0.2_mu0.1.npz"))
    field, psi, m, mu = d["field"], d["psi"], float(d["m"]), float(d["mu"])
    assert np.array_equal(oracle.fM(psi, field, m, mu), d["fM"])
    assert np.array_equal(oracle.fM(psi, field, m, mu, transpose=True), d["fMT"])
    assert np.array_equal(oracle.cg_MdM(psi, field, m, mu, propagator=True)[0], d["prop"])


@needs_ref_b
@pytest.mark.parametrize("nt,nx", [(16, 16), (16, 32), (32, 32)])
@pytest.mark.parametrize("mu,occ", [(0.0, 0.0), (0.1, 0.0), (0.3, 0.2)])
def test_flat_array_family_oracle_bitwise_vs_compiled_vec_ops(oracle, nt, nx, mu, occ):
    """vec_ops.c:345-461 (fM_occupied, fM_occupied_sq, action, cg_MdM_occupied; no caller in the reference)."""
    import ctypes as C

    rng = np.random.default_rng(nt * nx + int(100 * mu))
    ref = RefLibB(nt, nx, m=0.7, mu=mu)       # the mass must not enter (vec_ops.c:353 starts from chi = 0)
    field = (rng.random((nt, nx)) < occ).astype(np.int32)
    ref.set_field(field)
    psi = rng.normal(size=(nt, nx))
    assert np.array_equal(ref.call_flat("fM_occupied", psi), oracle.fM_occupied(psi, field, mu))
    assert np.array_equal(ref.call_flat("fM_occupied_sq", psi), oracle.fM_occupied(psi, field, mu, sq=True))
    ref.lib.action.restype = C.c_double
    ref.lib.action.argtypes = [C.c_void_p]
    assert ref.lib.action(psi.ctypes.data) == oracle.action(psi)
    # F F is negative definite on the free sites and the identity on occupied ones: sources supported on the free
    # sites keep the Krylov space inside one definite block
    src = np.where(field == 0, psi, 0.0)
    x_ref, ret_ref = ref.call_flat("cg_MdM_occupied", src)
    x, ret, it = oracle.cg_MdM_occupied(src, field, mu)
    assert ret == ret_ref
    if ret == 0:
        assert np.array_equal(x, x_ref)
        # psi = F (F F)^-1 source  =>  F psi = source
        assert np.abs(oracle.fM_occupied(x, field, mu) - src).max() < 1e-10
    # zero source: a = 0/0, the NaN branch returns 1 (vec_ops.c:448-451)
    assert ref.call_flat("cg_MdM_occupied", np.zeros((nt, nx)))[1] == 1
    assert oracle.cg_MdM_occupied(np.zeros((nt, nx)), field, mu)[1] == 1


@needs_ref_b
def test_flat_array_cg_with_occupied_dimers_bitwise(oracle):
    """Occupied nearest-neighbour pairs (what a fermion bag occupies) keep the two sublattices balanced: no zero
    modes, cg_MdM_occupied converges at mu = 0."""
    nt = nx = 32
    rng = np.random.default_rng(4)
    field = np.zeros((nt, nx), dtype=np.int32)
    for _ in range(10):
        t, x = rng.integers(nt), rng.integers(nx - 1)
        field[t, x] = field[t, x + 1] = 1
    ref = RefLibB(nt, nx, m=0.3, mu=0.0)
    ref.set_field(field)
    src = np.where(field == 0, rng.normal(size=(nt, nx)), 0.0)
    x_ref, ret_ref = ref.call_flat("cg_MdM_occupied", src)
    x, ret, it = oracle.cg_MdM_occupied(src, field, 0.0)
    assert ret == ret_ref == 0 and it > 50
    assert np.array_equal(x, x_ref)
