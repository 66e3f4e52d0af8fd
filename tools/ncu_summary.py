"""Summarise ncu outputs into the small text files committed under profiles/.

    python tools/ncu_summary.py launches gpurun_out/launches_X.csv  > profiles/launches_X.txt
    python tools/ncu_summary.py full gpurun_out/prof_X.ncu-rep      > profiles/ncu_full_X.txt
"""
import collections
import csv
import io
import re
import subprocess
import sys


def launches(path):
    lines = open(path).read().splitlines()
    start = next(i for i, l in enumerate(lines) if l.startswith('"ID"'))
    rows = list(csv.DictReader(io.StringIO("\n".join(lines[start:]))))
    agg = collections.OrderedDict()
    for r in rows:
        k = re.sub(r"\(.*", "", r["Kernel Name"]).replace("void <unnamed>::", "")
        a = agg.setdefault(k, [0, 0.0, r["Grid Size"], r["Block Size"]])
        a[0] += 1
        a[1] += float(r["Metric Value"])
    tot = sum(v[1] for v in agg.values())
    print(f"# {path}: {len(rows)} launches, {tot/1e3:.1f} us total (cold-cache, serialised: compare SHARES)")
    print(f"{'kernel':48s} {'n':>5s} {'avg_us':>9s} {'share':>7s}  grid / block")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{k:48s} {v[0]:5d} {v[1]/v[0]/1e3:9.2f} {v[1]/tot:7.3f}  {v[2]} / {v[3]}")


WANT = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
    "l1tex__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
    "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
    "smsp__inst_executed_pipe_fp64.sum", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__cycles_active.avg",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__cycles_elapsed.max",
    "smsp__average_warp_latency_issue_stalled_barrier.ratio", "smsp__warp_issue_stalled_barrier_per_warp_active.pct",
    "smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct",
    "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct",
    "smsp__warp_issue_stalled_mio_throttle_per_warp_active.pct",
    "smsp__warp_issue_stalled_math_pipe_throttle_per_warp_active.pct",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
]


def full(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    print(f"# {path}: ncu --set full, per launch")
    for r in rows[2:]:
        print("\n== " + r[hdr.index("Kernel Name")][:110])
        for w in WANT:
            if w in hdr:
                i = hdr.index(w)
                print(f"  {w:72s} {r[i]:>16s} {units[i]}")


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2])
