"""Generate the committed golden fixtures from the reference itself.

Run in the build container (needs oracle/_ref, i.e. /root/reference compiled by oracle/build_ref.sh):

    python tests/golden/make_golden.py

For every (lattice, flavour, m, mu) case the unmodified reference (hmc.c built as libhmcref_*.so) produces:
  A      gauge field after 5 heat-bath sweeps from the seeded Mersenne stream (update_puregauge_hb, hmc.c:82-93)
  v      stochastic_vector (hmc.c:439-447)
  Mv     fm_mul(v)                  (hmc.c:123-184)
  Mcv    fm_conjugate_mul(v)        (hmc.c:188-249; corrected hunk in the "adjoint" flavour)
  x      fmdm_invert_cg(Mcv)        (hmc.c:341-404), only where the reference converges
  dense_row0, dense_col0  first row / column of fermion_matrix() (hmc.c:269-310)
plus the stdout of the shipped `parameter` run (5 trajectories) as a text fixture.
"""
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.pyoracle import REF_DIR, RefLib  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
SEED = 4354365264  # the shipped parameter file's seed

CASES = [
    # nt, nx, flavour, m, g, mu, solve
    (8, 8, "compat", 100.0, 0.3, 0.1, True),
    (8, 8, "adjoint", 0.5, 0.3, 0.0, True),
    (16, 16, "compat", 100.0, 0.3, 0.1, True),
    (16, 16, "adjoint", 0.1, 0.3, 0.1, True),
    (16, 32, "adjoint", 1.0, 1.0, 0.3, True),
    (16, 32, "compat", 0.1, 1.0, 0.0, False),  # the reference exit(1)s here: apply only
    (32, 32, "compat", 100.0, 0.3, 0.1, True),
    (32, 32, "adjoint", 0.1, 0.3, 0.0, True),
    (64, 64, "adjoint", 0.3, 1.0, 0.05, True),
]


def main():
    for nt, nx, fl, m, g, mu, solve in CASES:
        ref = RefLib(nt, nx, fl, m=m, g=g, mu=mu, seed=SEED)
        G = ref.gauge()
        ref.heatbath(G, 5)
        v = ref.stochastic_vector()
        Mv = ref.fm_mul(v, G)
        Mcv = ref.fm_conjugate_mul(v, G)
        data = dict(A=G.A.copy(), v=v, Mv=Mv, Mcv=Mcv, m=m, g=g, mu=mu, mode=ref.mode)
        if solve:
            data["x"] = ref.fmdm_invert_cg(Mcv, G)
        if nt * nx <= 1024:
            D = ref.fermion_matrix(G)
            data["dense_row0"] = D[0].copy()
            data["dense_col0"] = D[:, 0].copy()
            data["dense_Dv"] = D @ v.ravel()
        name = f"ref_{nt}x{nx}_{fl}_m{m:g}_mu{mu:g}.npz"
        np.savez_compressed(os.path.join(OUT, name), **data)
        print("wrote", name)
    # stdout of the shipped parameter file (1000 1 100 0.3 0.1 4354365264) cut to 5 trajectories
    params = "5\n1\n100\n0.3\n0.1\n4354365264\n"
    out = subprocess.run([os.path.join(REF_DIR, "ref_hmc"), os.path.join(REF_DIR, "libhmcref_32x32_compat.so")],
                         input=params, capture_output=True, text=True, check=True).stdout
    with open(os.path.join(OUT, "hmc_32x32_shipped_5traj.stdout"), "w") as f:
        f.write(out)
    print("wrote hmc_32x32_shipped_5traj.stdout")
    # ADJOINT flavour, light-ish mass, measurement disabled (test_conjugate as coded aborts a true adjoint)
    params = "4\n100\n0.5\n0.3\n0.0\n4354365264\n"
    out = subprocess.run([os.path.join(REF_DIR, "ref_hmc"), os.path.join(REF_DIR, "libhmcref_16x16_adjoint.so")],
                         input=params, capture_output=True, text=True, check=True).stdout
    with open(os.path.join(OUT, "hmc_16x16_adjoint_m0.5_4traj.stdout"), "w") as f:
        f.write(out)
    print("wrote hmc_16x16_adjoint_m0.5_4traj.stdout")


def family_b():
    """vec_ops.c: fM, fM_transpose, cg_propagator on a 32x32 lattice with 15 % occupied sites."""
    from oracle.pyoracle import RefLibB
    rng = np.random.default_rng(2024)
    nt = nx = 32
    m, mu = 0.2, 0.1
    ref = RefLibB(nt, nx, m=m, mu=mu)
    field = (rng.random((nt, nx)) < 0.15).astype(np.int32)
    ref.set_field(field)
    psi = rng.normal(size=(nt, nx))
    np.savez_compressed(os.path.join(OUT, "refB_32x32_m0.2_mu0.1.npz"), field=field, psi=psi, m=m, mu=mu,
                        fM=ref.call("fM", psi), fMT=ref.call("fM_transpose", psi),
                        prop=ref.call("cg_propagator", psi))
    print("wrote refB_32x32_m0.2_mu0.1.npz")


if __name__ == "__main__":
    family_b()
    main()
