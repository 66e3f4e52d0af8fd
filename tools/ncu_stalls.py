"""Warp-stall breakdown of one kernel from an ncu report (no GPU needed):  python tools/ncu_stalls.py report.ncu-rep
Prints the duration, the FP64-pipe and issue utilisation and the sampled stall reasons as shares of all samples."""
import csv
import io
import subprocess
import sys


def main():
    rep = sys.argv[1]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr = rows[0]
    for vals in rows[2:]:
        d = dict(zip(hdr, vals))
        print(d.get("Kernel Name", "?")[:110])
        for k in ("gpu__time_duration.sum", "launch__grid_size", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
                  "smsp__issue_active.avg.pct_of_peak_sustained_active", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"):
            print(f"  {k} = {d.get(k)}")
        st = {k[len("smsp__pcsamp_warps_issue_stalled_"):]: float(v.replace(",", "")) for k, v in d.items()
              if k.startswith("smsp__pcsamp_warps_issue_stalled_") and "not_issued" not in k and v}
        tot = sum(st.values())
        print("  stalls: " + ", ".join(f"{k} {100 * v / tot:.1f}%" for k, v in sorted(st.items(), key=lambda kv: -kv[1]) if v / tot > 0.005))


if __name__ == "__main__":
    main()
