"""T6: the reference's UNMODIFIED hmc.c driver running on top of the GPU library through symbol interposition
 reproduces the reference's own stdout.  The driver shared objects are the prebuilt
oracle/_ref/libhmcref_*.so (the reference compiled as-is); only the checker side uses them."""
import os
import subprocess

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")
REF = os.path.join(ROOT, "oracle", "_ref")
LAUNCHER = os.path.join(ROOT, "thirring2d_b200", "hmc_b200")


def run_interposed(lib, nt, nx, mode, params, coarse=False, env=None):
    so = os.path.join(REF, lib)
    if not os.path.exists(so) or not os.path.exists(LAUNCHER):
        pytest.skip("prebuilt reference driver or launcher missing")
    cmd = [LAUNCHER, so, str(nt), str(nx), mode] + (["0", "coarse"] if coarse else [])
    p = subprocess.run(cmd, input=params, capture_output=True, text=True, timeout=600,
                       env=dict(os.environ, **(env or {})))
    assert p.returncode == 0, p.stderr
    return p.stdout, p.stderr


def test_T6_shipped_parameter_file_stdout_is_reproduced():
    """32x32, m=100, g=0.3, mu=0.1, seed 4354365264 (the shipped `parameter` file, first 5 trajectories):
    every printed line (6 significant digits) equals the CPU reference's."""
    out, err = run_interposed("libhmcref_32x32_compat.so", 32, 32, "compat", "5\n1\n100\n0.3\n0.1\n4354365264\n")
    gold = open(os.path.join(GOLD, "hmc_32x32_shipped_5traj.stdout")).read()
    assert out == gold
    # 11 CG solves per trajectory (hmc.c:515,719) and the applies of measure() all went through the GPU
    assert "55 CG solves" in err, err


def test_T6_adjoint_mode_driver_matches_corrected_reference():
    out, err = run_interposed("libhmcref_16x16_adjoint.so", 16, 16, "adjoint", "4\n100\n0.5\n0.3\n0.0\n4354365264\n")
    gold = open(os.path.join(GOLD, "hmc_16x16_adjoint_m0.5_4traj.stdout")).read()
    assert out == gold


def test_T5_interposed_divergence_prints_and_exits_like_the_reference():
    """REF_COMPAT at m = 0.1: 'Cannot invert fermion matrix' + exit(1) (hmc.c:383-388)."""
    so = os.path.join(REF, "libhmcref_16x16_compat.so")
    if not os.path.exists(so):
        pytest.skip("prebuilt reference driver missing")
    p = subprocess.run([LAUNCHER, so, "16", "16", "compat"], input="2\n1\n0.1\n0.3\n0.0\n4354365264\n",
                       capture_output=True, text=True, timeout=600)
    assert p.returncode == 1
    assert "Cannot invert fermion matrix" in p.stdout


def test_coarse_override_update_gauge_is_one_device_trajectory():
    """SURVEY 8(b) optional coarse override: libthirring_hmc_coarse.so in front of the unmodified driver turns
    update_gauge (hmc.c:671-746) into one device-resident trajectory fed by the driver's own Mersenne stream.  The
    stdout of the shipped parameter file is still the reference's, byte for byte."""
    out, err = run_interposed("libhmcref_32x32_compat.so", 32, 32, "compat", "5\n1\n100\n0.3\n0.1\n4354365264\n",
                              coarse=True)
    gold = open(os.path.join(GOLD, "hmc_32x32_shipped_5traj.stdout")).read()
    assert out == gold
    assert "5 whole trajectories" in err and " 0 CG solves" in err, err


def test_coarse_override_adjoint_and_light_mass():
    out, err = run_interposed("libhmcref_16x16_adjoint.so", 16, 16, "adjoint", "4\n100\n0.5\n0.3\n0.0\n4354365264\n",
                              coarse=True)
    gold = open(os.path.join(GOLD, "hmc_16x16_adjoint_m0.5_4traj.stdout")).read()
    assert out == gold
    # 40 leapfrog steps at m = 0.1: the coarse path against the fine-grained path of the same library
    so = "libhmcref_32x32_adjoint_ns40.so"
    par = "3\n100\n0.1\n0.3\n0.0\n4354365264\n"
    fine, _ = run_interposed(so, 32, 32, "adjoint", par)
    coarse, err = run_interposed(so, 32, 32, "adjoint", par, coarse=True, env={"THIRRING_NSTEPS": "40"})
    assert coarse == fine
    assert "3 whole trajectories" in err


def test_coarse_override_divergence_prints_and_exits_like_the_reference():
    so = os.path.join(REF, "libhmcref_16x16_compat.so")
    if not os.path.exists(so):
        pytest.skip("prebuilt reference driver missing")
    p = subprocess.run([LAUNCHER, so, "16", "16", "compat", "0", "coarse"], input="2\n1\n0.1\n0.3\n0.0\n4354365264\n",
                       capture_output=True, text=True, timeout=600)
    assert p.returncode == 1
    assert "Cannot invert fermion matrix" in p.stdout
