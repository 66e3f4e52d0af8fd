"""Host logic of the batched driver: the stdout line format equals the reference's printf lines (hmc.c:701,735,
739,743,839-840) up to the chain prefix, checked against the recorded reference stdout.  CPU only."""
import io
import os
import re

import numpy as np

from thirring2d_b200.hmc_driver import banner, measurement_lines, read_parameters, trajectory_lines

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_parameter_file_is_read_like_the_reference():
    shipped = "1000\n1\n100\n0.3\n0.1\n4354365264\n"   # /root/reference/parameter
    assert read_parameters(io.StringIO(shipped)) == (1000, 1, 100.0, 0.3, 0.1, 4354365264)


def test_lines_round_trip_the_reference_stdout():
    out = open(os.path.join(GOLD, "hmc_32x32_shipped_5traj.stdout")).read()
    starts = re.findall(r"Start HMC: Sg (\S+), Smdm (\S+), Smd (\S+), Smom (\S+)", out)
    ends = re.findall(r"HMC End, dS (\S+), Sg (\S+), Smdm (\S+), Smd (\S+), Sm (\S+)", out)
    accs = re.findall(r"HMC (ACCEPTED|REJECTED)", out)
    mags = re.findall(r"Magnetisation (\S+)", out)
    phs = re.findall(r"Phase (\S+)", out)
    assert len(starts) == len(ends) == len(accs) == len(mags) == 5
    ref_lines = [l for l in out.splitlines() if re.match(r"(Start HMC|HMC |Magnetisation|Phase)", l)]
    mine = []
    for s, e, a, mg, ph in zip(starts, ends, accs, mags, phs):
        obs = [float(v) for v in s] + [float(v) for v in e[1:]] + [float(e[0]), 1.0 if a == "ACCEPTED" else 0.0]
        mine += trajectory_lines(obs, 7) + measurement_lines(float(mg), float(ph), 7)
    assert [l.replace("[chain 7] ", "") for l in mine] == ref_lines
    assert all(l.startswith("[chain 7] ") for l in mine)


def test_banner_matches_reference():
    out = open(os.path.join(GOLD, "hmc_32x32_shipped_5traj.stdout")).read()
    for line in banner(32, 32, 1, 100.0, 0.3, 0.1, 4354365264)[1:]:
        assert line in out, line


def test_scan_parsing_and_summary_lines():
    from thirring2d_b200.hmc_driver import parse_scan, summary_lines

    pts = parse_scan("0.2,0.4:0.01,0.1,0.3")
    assert pts == [(0.2, 0.01), (0.2, 0.1), (0.2, 0.3), (0.4, 0.01), (0.4, 0.1), (0.4, 0.3)]
    cnt = np.array([4.0] * 6)
    mean = np.arange(12, dtype=float).reshape(6, 2)
    err = np.full((6, 2), 0.5)
    lines = summary_lines(pts, cnt, mean, err, ["acceptance", "Condensate"])
    assert lines[0] == "[point 0 g 0.2 m 0.01] chains 4, acceptance 0 +- 0.5, Condensate 1 +- 0.5"
    assert len(lines) == 6 and lines[5].startswith("[point 5 g 0.4 m 0.3] chains 4, acceptance 10 +- 0.5")
