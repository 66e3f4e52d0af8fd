"""Device-resident batched trajectory against the reference's own update_gauge / measure
(hmc.c:671-746, 794-842) fed with the SAME random numbers: the reference's Mersenne stream is replayed on the
host, turned into the Box-Muller fields the reference would build, and handed to tb_hmc_trajectory."""
import ctypes
import os
import re

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

tb = pytest.importorskip("thirring2d_b200")
from oracle.pyoracle import RefLib, ref_available  # noqa: E402  (the checker)

SEED = 4354365264


def box_muller_fields(ref, n, nt, nx):
    """n vectors of (sqrt(-2 ln x1) cos 2 pi x2) + i (... sin 2 pi x2), x1 then x2 per site in (t,x) order
    (hmc.c:422-426, 441-446, 486-491)."""
    out = np.empty((n, nt, nx), dtype=np.complex128)
    for k in range(n):
        draws = np.array([ref.mersenne() for _ in range(2 * nt * nx)]).reshape(nt, nx, 2)
        rad = np.sqrt(-2.0 * np.log(draws[..., 0]))
        out[k] = rad * np.cos(2 * np.pi * draws[..., 1]) + 1j * rad * np.sin(2 * np.pi * draws[..., 1])
    return out


def make_refs(nt, nx, flavour, m, g, mu, sweeps=5, nsteps=10):
    """Two identically seeded copies of the reference: one to run update_gauge, one to replay its draws."""
    refs = []
    for _ in range(2):
        r = RefLib(nt, nx, flavour, m=m, g=g, mu=mu, seed=SEED, nsteps=nsteps)
        G = r.gauge()
        r.heatbath(G, sweeps)
        refs.append((r, G))
    return refs


CASES = [
    (16, 16, "adjoint", 0.5, 0.3, 0.0, 10),
    (32, 32, "compat", 100.0, 0.3, 0.1, 10),   # the shipped parameter file
    (32, 32, "adjoint", 0.1, 0.3, 0.0, 40),    # light mass needs the 40-step oracle copy
]


@pytest.mark.parametrize("nt,nx,flavour,m,g,mu,nsteps", CASES)
def test_trajectory_matches_reference_update_gauge(capfd, nt, nx, flavour, m, g, mu, nsteps):
    if not ref_available(nt, nx, flavour, nsteps):
        pytest.skip("oracle/_ref not built")
    (ref, G), (rep, G2) = make_refs(nt, nx, flavour, m, g, mu, nsteps=nsteps)
    libc = ctypes.CDLL(None)
    mode = tb.MODE_ADJOINT if flavour == "adjoint" else tb.MODE_REF_COMPAT
    with tb.Context(nt, nx, 1, mode, m=m, mu=mu) as ctx:
        ctx.hmc_set_coupling(g)
        for traj in range(3):
            A0 = G.A.copy()
            assert np.array_equal(A0, G2.A)
            # replay the draws update_gauge is about to consume: pseudofermion, momentum, stochastic vector, accept
            xi, pm, st = box_muller_fields(rep, 3, nt, nx)
            u = rep.mersenne()
            mom = np.stack([pm.real, pm.imag], axis=-1)
            capfd.readouterr()
            ref.lib.update_gauge(G.top.ctypes.data)
            libc.fflush(None)
            out = capfd.readouterr().out
            G2.arr[...] = G.arr  # keep the replay copy's field in step
            start = [float(v) for v in re.search(r"Start HMC: Sg (\S+), Smdm (\S+), Smd (\S+), Smom (\S+)", out).groups()]
            end = [float(v) for v in
                   re.search(r"HMC End, dS (\S+), Sg (\S+), Smdm (\S+), Smd (\S+), Sm (\S+)", out).groups()]
            accepted_ref = "HMC ACCEPTED" in out

            ctx.set_gauge(A0)
            obs, acc, iters = ctx.hmc_trajectory(nsteps=nsteps, traj_length=1.0, xi=xi[None], mom=mom[None], st=st[None],
                                                 u=np.array([u]))
            o = obs[0]
            # printed with %g: 6 significant digits
            for got, want in zip(o[0:4], start):
                assert abs(got - want) <= 6e-6 * abs(want), (traj, o, start)
            for got, want in zip(o[4:8], end[1:]):
                assert abs(got - want) <= 6e-6 * abs(want), (traj, o, end)
            scale = max(abs(v) for v in start)
            assert abs(o[8] - end[0]) <= 6e-6 * abs(end[0]) + 1e-9 * scale, (traj, o[8], end[0])
            assert bool(acc[0]) == accepted_ref
            A_gpu = ctx.get_gauge()[0]
            assert np.allclose(A_gpu, G.A, rtol=0, atol=1e-9), np.abs(A_gpu - G.A).max()
            assert iters > 0


def test_measure_matches_reference(capfd):
    """Magnetisation and Phase of measure() (hmc.c:823-842) with the reference's own 20 sources."""
    nt = nx = 16
    if not ref_available(nt, nx, "compat"):
        pytest.skip("oracle/_ref not built")
    (ref, G), (rep, G2) = make_refs(nt, nx, "compat", 100.0, 0.3, 0.1)
    libc = ctypes.CDLL(None)
    ctypes.c_void_p.in_dll(ref.lib, "A").value = G.top.ctypes.data   # measure() reads the global A
    # draws: test_conjugate consumes one stochastic vector first (hmc.c:764), then 20 phase sources (hmc.c:801-803)
    box_muller_fields(rep, 1, nt, nx)
    src = box_muller_fields(rep, 20, nt, nx)
    capfd.readouterr()
    ref.lib.measure()
    libc.fflush(None)
    out = capfd.readouterr().out
    mag_ref = float(re.search(r"Magnetisation (\S+)", out).group(1))
    ph_ref = float(re.search(r"Phase (\S+)", out).group(1))
    with tb.Context(nt, nx, 1, tb.MODE_REF_COMPAT, m=100.0, mu=0.1) as ctx:
        ctx.set_gauge(G.A)
        mag, ph = ctx.hmc_measure(nsrc=20, sources=src[:, None])
    assert abs(mag[0] - mag_ref) <= 6e-6 * abs(mag_ref)
    assert abs(ph[0] - ph_ref) <= 6e-6 * abs(ph_ref) + 1e-9


def test_batched_trajectories_device_rng_statistics():
    """256 chains with the device Philox stream: sane acceptance, <exp(-dS)> and reversibility-independent checks.
    Statistical parity only (the chain as coded has <exp(-dS)> != 1, so only loose bounds)."""
    nt = nx = 16
    n = 128
    with tb.Context(nt, nx, n, tb.MODE_ADJOINT, m=0.5, mu=0.0) as ctx:
        ctx.hmc_set_coupling(0.3)
        ctx.hmc_heatbath(100, seed=5)
        A0 = ctx.get_gauge()
        # equilibrium of the quenched heat bath: P(A) ~ exp((Nf/g) cos A) => <cos A> = I1(k)/I0(k), k = 2/0.3
        k = 2 / 0.3
        from scipy.special import i0, i1
        assert abs(np.cos(A0).mean() - i1(k) / i0(k)) < 5e-3
        obs, acc, iters = ctx.hmc_trajectory(nsteps=20, traj_length=0.5, seed=9, traj_index=0)
        assert np.all(np.isfinite(obs))
        assert 0.3 < acc.mean() <= 1.0
        # start-of-trajectory actions have their heat-bath expectation values: Smdm = |xi|^2 ~ 2V, Smom ~ 2V
        V = nt * nx
        assert abs(obs[:, 1].mean() / (2 * V) - 1) < 0.03 and abs(obs[:, 3].mean() / (2 * V) - 1) < 0.03
        A1 = ctx.get_gauge()
        changed = np.abs(A1 - A0).reshape(n, -1).max(axis=1) > 0
        assert np.array_equal(changed, acc.astype(bool))  # exactly the accepted chains moved
        obs2, acc2, _ = ctx.hmc_trajectory(nsteps=20, traj_length=0.5, seed=9, traj_index=1)
        assert not np.array_equal(obs[:, 1], obs2[:, 1])  # a new trajectory index draws new fields


def test_checkpoint_round_trip_and_batched_driver(tmp_path):
    """On-disk format: raw FP64 gauge dump with a header, and the batched driver printing the
    reference's stdout keywords per chain; a resumed run continues from the stored fields."""
    import subprocess
    import sys

    nt = nx = 16
    ck = str(tmp_path / "gauge.ckpt")
    with tb.Context(nt, nx, 6, tb.MODE_ADJOINT, m=0.5) as ctx:
        ctx.hmc_set_coupling(0.3)
        ctx.hmc_heatbath(20, seed=3)
        A = ctx.get_gauge()
        ctx.checkpoint_write(ck)
    assert os.path.getsize(ck) == 64 + A.nbytes
    with tb.Context(nt, nx, 6, tb.MODE_ADJOINT, m=0.5) as ctx:
        ctx.checkpoint_read(ck)
        assert np.array_equal(ctx.get_gauge(), A)
    with tb.Context(nt, nx, 5, tb.MODE_ADJOINT, m=0.5) as ctx:
        with pytest.raises(tb.TBError, match="holds 6 chains"):
            ctx.checkpoint_read(ck)
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    p = subprocess.run([sys.executable, "-m", "thirring2d_b200.hmc_driver", "--nt", "16", "--nx", "16", "--chains", "6",
                        "--mode", "adjoint", "--nsteps", "10", "--resume", ck], input="2\n2\n0.5\n0.3\n0.0\n99\n",
                       capture_output=True, text=True, cwd=root, timeout=300)
    assert p.returncode == 0, p.stderr
    out = p.stdout
    assert " 2D quenched Thirring model, ( 16 , 16 ) lattice" in out
    assert len(re.findall(r"^\[chain \d\] Start HMC: Sg \S+, Smdm \S+, Smd \S+, Smom \S+$", out, re.M)) == 12
    assert len(re.findall(r"^\[chain \d\] HMC End, dS \S+, Sg \S+, Smdm \S+, Smd \S+, Sm \S+$", out, re.M)) == 12
    assert len(re.findall(r"^\[chain \d\] HMC (ACCEPTED|REJECTED)$", out, re.M)) == 12
    assert len(re.findall(r"^\[chain \d\] Magnetisation \S+$", out, re.M)) == 6
    assert len(re.findall(r"^\[chain \d\] Phase \S+$", out, re.M)) == 6
    # the first printed gauge action is the action of the checkpointed field: (Nf/g) sum (1 - cos A)
    sg0 = float(re.search(r"^\[chain 0\] Start HMC: Sg (\S+),", out, re.M).group(1))
    assert abs(sg0 - (2 / 0.3) * np.sum(1 - np.cos(A[0]))) <= 6e-6 * sg0
