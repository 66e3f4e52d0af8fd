"""T8 golden record: independent Markov chains of the REFERENCE's own driver functions (update_puregauge_hb and
update_gauge of the compiled, ADJOINT-corrected hmc.c, oracle/_ref/libhmcref_32x32_adjoint.so) with different Mersenne
seeds, observables parsed from the reference's own stdout at every trajectory.  The GPU test compares its device-RNG
chains with these at equal trajectory index (the run is not thermalised).  Computing them takes minutes of CPU, so the
outcome is committed:  tests/golden/ensemble_32x32_m0.5_g0.3.npz  [chain][trajectory][Sg at start, dS, accepted, Magnetisation]

    python tests/golden/make_golden_ensemble.py          (about 3 minutes)
"""
import ctypes
import os
import re
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.pyoracle import RefLib  # noqa: E402

NT = NX = 32
M, G, MU, NSTEPS, NTRAJ, SWEEPS, NCHAINS = 0.5, 0.3, 0.0, 10, 10, 20, 96


def capture_stdout(fn):
    """Run fn() and return what the C library printed to file descriptor 1."""
    libc = ctypes.CDLL(None)
    sys.stdout.flush()
    with tempfile.TemporaryFile(mode="w+b") as tmp:
        saved = os.dup(1)
        os.dup2(tmp.fileno(), 1)
        try:
            fn()
            libc.fflush(None)
        finally:
            os.dup2(saved, 1)
            os.close(saved)
        tmp.seek(0)
        return tmp.read().decode()


def main():
    obs = np.zeros((NCHAINS, NTRAJ, 4))
    for c in range(NCHAINS):
        r = RefLib(NT, NX, "adjoint", m=M, g=G, mu=MU, nsteps=NSTEPS)
        r.seed(1000 + 7 * c, warmup=2000)
        Gf = r.gauge()
        r.heatbath(Gf, SWEEPS)
        for t in range(NTRAJ):
            out = capture_stdout(lambda: r.lib.update_gauge(Gf.top.ctypes.data))
            sg = float(re.search(r"Start HMC: Sg (\S+),", out).group(1))
            ds = float(re.search(r"HMC End, dS (\S+),", out).group(1))
            obs[c, t] = sg, ds, float("HMC ACCEPTED" in out), Gf.A.sum() / (NT * NX)
        print(f"chain {c}: acceptance {obs[c, :, 2].mean():.2f}", file=sys.stderr)
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ensemble_32x32_m0.5_g0.3.npz")
    np.savez(out, obs=obs, nt=NT, nx=NX, m=M, g=G, mu=MU, nsteps=NSTEPS, sweeps=SWEEPS,
             columns=np.array(["Sg_start", "dS", "accepted", "Magnetisation"]))
    print("acceptance", obs[:, :, 2].mean(), "mean dS", np.median(obs[:, :, 1]))


if __name__ == "__main__":
    main()
