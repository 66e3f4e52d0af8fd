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
        ctx.checkpoint_write(ck, next_trajectory=3)
    assert os.path.getsize(ck) == 64 + A.nbytes
    with tb.Context(nt, nx, 6, tb.MODE_ADJOINT, m=0.5) as ctx:
        assert ctx.checkpoint_read(ck) == 3   # the trajectory counter travels in the header
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


def test_resumed_run_continues_the_random_stream(tmp_path):
    """A run of 4 trajectories and a run of 2 + a resumed run of 2 print the same trajectories 3 and 4: the checkpoint
    carries the trajectory index that keys the device random stream, so the second leg does not replay the momenta,
    noise and Metropolis uniforms of the first.  A file without the index is refused unless --traj-offset is given."""
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    base = [sys.executable, "-m", "thirring2d_b200.hmc_driver", "--nt", "16", "--nx", "16", "--chains", "3",
            "--mode", "adjoint", "--nsteps", "10", "--condensate", "2"]

    def run(extra, n_loops):
        return subprocess.run(base + extra, input=f"{n_loops}\n1\n0.5\n0.3\n0.0\n77\n", capture_output=True, text=True,
                              cwd=root, timeout=300)

    ck = str(tmp_path / "leg1.ckpt")
    full = run([], 4)
    leg1 = run(["--checkpoint", ck], 2)
    leg2 = run(["--resume", ck], 2)
    assert full.returncode == 0 and leg1.returncode == 0 and leg2.returncode == 0, (full.stderr, leg1.stderr, leg2.stderr)

    def chain_lines(out):
        return [ln for ln in out.splitlines() if ln.startswith("[chain")]

    per_traj = len(chain_lines(full.stdout)) // 4
    assert chain_lines(leg1.stdout) == chain_lines(full.stdout)[:2 * per_traj]
    assert chain_lines(leg2.stdout) == chain_lines(full.stdout)[2 * per_traj:]
    # a checkpoint that does not record the index (written by the bare C call without the counter set)
    old = str(tmp_path / "old.ckpt")
    with tb.Context(16, 16, 3, tb.MODE_ADJOINT, m=0.5) as ctx:
        ctx.checkpoint_read(ck)
        ctx.checkpoint_write(old, next_trajectory=0)
    refused = run(["--resume", old], 1)
    assert refused.returncode != 0 and "--traj-offset" in (refused.stderr + refused.stdout)
    ok = run(["--resume", old, "--traj-offset", "3"], 2)
    assert ok.returncode == 0 and chain_lines(ok.stdout) == chain_lines(leg2.stdout)


def test_force_is_the_derivative_of_the_action_check_force():
    """The reference's CHECK_FORCE (hmc.c:502,535-559, disabled in the shipped source) on the GPU, batched: the force
    momentum_step applies must be the derivative of Sg + Re<psi, (M^dagger M)^-1 psi> with respect to every link angle.
    One chain per probed link and sign (A +- h e_k): 2 x 96 perturbed fields solved as one batch; central differences
    with h = 1e-4 against tb_hmc_force on the unperturbed field.  ADJOINT mode (the formula assumes the true M^dagger),
    mu != 0, the wrap-around links included."""
    nt = nx = 16
    m, mu, g, h = 0.4, 0.15, 0.3, 1e-4
    rng = np.random.default_rng(31)
    A = rng.uniform(-np.pi, np.pi, size=(nt, nx, 2))
    psi = rng.normal(size=(nt, nx)) + 1j * rng.normal(size=(nt, nx))
    links = [(0, 0, 0), (nt - 1, 3, 0), (5, nx - 1, 1), (nt - 1, nx - 1, 1), (0, nx - 1, 0), (7, 0, 1)]
    links += [(int(t), int(x), int(d)) for t, x, d in zip(rng.integers(0, nt, 90), rng.integers(0, nx, 90), rng.integers(0, 2, 90))]
    n = 2 * len(links)
    Ab = np.broadcast_to(A, (n, nt, nx, 2)).copy()
    for k, (t, x, d) in enumerate(links):
        Ab[2 * k, t, x, d] += h
        Ab[2 * k + 1, t, x, d] -= h
    with tb.Context(nt, nx, n, tb.MODE_ADJOINT, m=m, mu=mu) as ctx:
        ctx.set_gauge(Ab)
        xs, info = ctx.fmdm_invert_cg(np.broadcast_to(psi, (n, nt, nx)))
        assert np.all(info.status == tb.CG_CONVERGED)
    S = (2.0 / g) * (1.0 - np.cos(Ab)).reshape(n, -1).sum(axis=1) + np.array([np.vdot(psi, xs[c]).real for c in range(n)])
    with tb.Context(nt, nx, 1, tb.MODE_ADJOINT, m=m, mu=mu) as ctx:
        ctx.hmc_set_coupling(g)
        ctx.set_gauge(A)
        F = ctx.hmc_force(psi)[0]
    for k, (t, x, d) in enumerate(links):
        fd = (S[2 * k] - S[2 * k + 1]) / (2 * h)
        assert abs(F[t, x, d] - fd) <= 1e-6 * max(1.0, abs(fd)), ((t, x, d), F[t, x, d], fd)


def free_field_condensate(L, m):
    """(1/V) Tr M^-1 at A = 0, mu = 0: (1/V) sum_k m / (m^2 + sum_mu sin^2 k_mu), k antiperiodic."""
    k = (2 * np.arange(L) + 1) * np.pi / L
    s = np.sin(k) ** 2
    return float((m / (m * m + s[:, None] + s[None, :])).mean())


@pytest.mark.parametrize("L,solver", [(16, 0), (8, 1)])
def test_condensate_free_field(L, solver):
    """Stochastic estimator against the closed form on the free field, both solvers (resident 16^2, streaming 8^2)."""
    m, n, nsrc = 0.5, 64, 20
    with tb.Context(L, L, n, tb.MODE_ADJOINT, m=m, mu=0.0) as ctx:
        ctx.set_tuning(solver=solver)
        ctx.set_gauge(np.zeros((n, L, L, 2)))
        cond, iters = ctx.hmc_condensate(nsrc=nsrc, seed=11)
    want = free_field_condensate(L, m)
    err = cond.std(ddof=1) / np.sqrt(n)
    assert iters > 0 and np.all(np.isfinite(cond))
    assert abs(cond.mean() - want) < 4 * err, (cond.mean(), want, err)
    assert err < 0.01 * want   # the estimator is not trivially noisy: 64 x 20 sources pin it to < 1 %


@pytest.mark.parametrize("flavour,m,mu", [("adjoint", 0.3, 0.0), ("adjoint", 0.5, 0.2), ("compat", 100.0, 0.1)])
def test_condensate_matches_reference_fm_invert_cg(flavour, m, mu):
    """Same sources through the reference's own fm_invert_cg (hmc.c:408-414) on a heat-bath field."""
    nt = nx = 16
    if not ref_available(nt, nx, flavour):
        pytest.skip("oracle/_ref not built")
    ref = RefLib(nt, nx, flavour, m=m, g=0.3, mu=mu, seed=SEED, nsteps=10)
    G = ref.gauge()
    ref.heatbath(G, 5)
    nsrc, n = 4, 3
    rng = np.random.default_rng(5)
    src = rng.standard_normal((nsrc, n, nt, nx)) + 1j * rng.standard_normal((nsrc, n, nt, nx))
    want = np.zeros(n)
    for i in range(nsrc):
        for c in range(n):
            want[c] += np.vdot(src[i, c], ref.fm_invert_cg(src[i, c], G)).real
    want /= 2 * nt * nx * nsrc
    mode = tb.MODE_ADJOINT if flavour == "adjoint" else tb.MODE_REF_COMPAT
    with tb.Context(nt, nx, n, mode, m=m, mu=mu) as ctx:
        ctx.set_gauge(np.broadcast_to(G.A, (n, nt, nx, 2)))
        cond, _ = ctx.hmc_condensate(nsrc=nsrc, sources=src)
    assert np.allclose(cond, want, rtol=1e-10, atol=0), (cond, want)


def test_ensemble_matches_reference_chains():
    """T8: 1 024 device-RNG chains on the GPU against 96 independent chains of the reference's own driver
    functions (update_puregauge_hb + update_gauge of the compiled hmc.c, different Mersenne seeds; committed as
    tests/golden/ensemble_32x32_m0.5_g0.3.npz by tests/golden/make_golden_ensemble.py), 32 x 32, m = 0.5, g = 0.3,
    10 leapfrog steps as the reference hard-codes, compared at equal trajectory index over 10 trajectories -- the run is
    not thermalised, so equal index, not 'equilibrium'.  Every (trajectory, observable) pair gives a z-score against the
    combined statistical error; each must be inside 4 and together they must look like unit Gaussians (3 sigma of the
    chi^2 of 40 scores).  dS enters through the bounded acceptance probability min(1, exp(-dS)): its own distribution
    has a heavy upper tail."""
    gold = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ensemble_32x32_m0.5_g0.3.npz"))
    ref = gold["obs"]                     # [chain][trajectory][Sg at start, dS, accepted, Magnetisation]
    nt, nx = int(gold["nt"]), int(gold["nx"])
    ntraj, n_gpu = ref.shape[1], 1024
    gpu = np.zeros((n_gpu, ntraj, 4))
    with tb.Context(nt, nx, n_gpu, tb.MODE_ADJOINT, m=float(gold["m"]), mu=float(gold["mu"])) as ctx:
        ctx.hmc_set_coupling(float(gold["g"]))
        ctx.hmc_heatbath(int(gold["sweeps"]), seed=77)
        for t in range(ntraj):
            obs, acc, _ = ctx.hmc_trajectory(nsteps=int(gold["nsteps"]), traj_length=1.0, seed=77, traj_index=t + 1)
            assert not ctx.hmc_cg_failures().any()
            mag, _ = ctx.hmc_measure(nsrc=0)
            gpu[:, t] = np.stack([obs[:, 0], obs[:, 8], acc, mag], axis=1)

    def columns(o):   # Sg, acceptance probability, accepted, Magnetisation
        return np.stack([o[..., 0], np.minimum(1.0, np.exp(-o[..., 1])), o[..., 2], o[..., 3]], axis=-1)

    a, b = columns(ref), columns(gpu)
    zs = []
    names = ["Sg", "P_acc", "accepted", "Magnetisation"]
    for t in range(ntraj):
        for k in range(4):
            err = np.sqrt(a[:, t, k].var(ddof=1) / a.shape[0] + b[:, t, k].var(ddof=1) / b.shape[0])
            if err == 0.0:   # e.g. the first trajectory from the heat-bath start: dS << 0, every chain accepts on both sides
                assert a[:, t, k].mean() == b[:, t, k].mean(), (t, names[k])
                continue
            zs.append(((a[:, t, k].mean() - b[:, t, k].mean()) / err, t, names[k]))
    z = np.array([v[0] for v in zs])
    worst = max(zs, key=lambda v: abs(v[0]))
    assert np.abs(z).max() < 4.0, worst
    chi2, n = float((z ** 2).sum()), z.size
    assert chi2 < n + 3.0 * np.sqrt(2.0 * n), (chi2, n, z.round(2).tolist())
    # the comparison has teeth: the gauge action is pinned to better than 1 %, the acceptance to 0.05
    err_sg = np.sqrt(a[:, 0, 0].var(ddof=1) / a.shape[0] + b[:, 0, 0].var(ddof=1) / b.shape[0])
    assert err_sg < 0.01 * a[:, 0, 0].mean()
    assert abs(a[..., 2].mean() - b[..., 2].mean()) < 0.05 and 0.5 < b[..., 2].mean() < 0.9


def _run_driver(args, stdin, nproc=1):
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    if nproc == 1:
        cmd = [sys.executable, "-m", "thirring2d_b200.hmc_driver"] + args
    else:
        import socket

        with socket.socket() as s:
            s.bind(("127.0.0.1", 0))
            port = s.getsockname()[1]
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}", "--master-addr",
               "127.0.0.1", "--master-port", str(port), "-m", "thirring2d_b200.hmc_driver"] + args
    p = subprocess.run(cmd, input=stdin, capture_output=True, text=True, cwd=root, timeout=600)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-4000:]
    return p.stdout


def test_scan_driver_is_independent_of_the_number_of_ranks():
    """Coupling/mass scan (BASELINE config 5) through the batched driver: 4 (g, m) points x 3 chains.  The device
    random stream is keyed by the GLOBAL chain index, so the same ensemble comes out whether one process owns all
    chains or two ranks own half each (here both ranks share the one GPU, gloo for the final reduction): the
    per-chain lines are identical and the summary agrees to the printed digits."""
    args = ["--nt", "16", "--nx", "16", "--chains", "3", "--mode", "adjoint", "--nsteps", "10", "--traj-length", "0.5",
            "--scan", "0.3,0.6:0.3,1.0", "--condensate", "3"]
    params = "2\n2\n0.5\n0.3\n0.0\n77\n"
    one = _run_driver(args, params)
    two = _run_driver(args + ["--backend", "gloo"], params, nproc=2)
    chain_lines = lambda out: sorted(l for l in out.splitlines() if l.startswith("[chain "))
    points = lambda out: [l for l in out.splitlines() if l.startswith("[point ")]
    assert len(chain_lines(one)) == 12 * (2 * 3 + 3)      # 2 trajectories x 3 lines + 3 measurement lines per chain
    assert chain_lines(one) == chain_lines(two)
    assert len(points(one)) == 4 and points(one) == points(two)
    assert points(one)[0].startswith("[point 0 g 0.3 m 0.3] chains 3, acceptance ")
    # heavier fermions -> smaller condensate ratio <psibar psi>(m=0.3) / <psibar psi>(m=1.0) > 1 is NOT generic, but the
    # free-field order of magnitude is: (1/V) Tr M^-1 lies between m/(m^2+2) and 1/m
    for l in points(one):
        m = float(re.search(r" m (\S+)\]", l).group(1))
        cond = float(re.search(r"Condensate (\S+) \+-", l).group(1))
        assert m / (m * m + 2.0) < cond < 1.0 / m, l
