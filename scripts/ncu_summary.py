"""Summarise an .ncu-rep (read on the CPU box): per kernel, the metrics that decide what bounds it."""
import csv
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "launch__registers_per_thread", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "smsp__warps_eligible.avg.per_cycle_active",
]


def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    for r in rows[2:]:
        print("-----", r[idx["Kernel Name"]][:70], r[idx["Grid Size"]], r[idx["Block Size"]])
        for k in KEYS:
            if k in idx:
                print(f"   {k:72s} {r[idx[k]]:>16s} {units[idx[k]]}")
        stalls = []
        for h in hdr:
            if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio"):
                try:
                    stalls.append((float(r[idx[h]]), h.split("stalled_")[1].split("_per_issue")[0]))
                except ValueError:
                    pass
        print("   stalls (warps per issue):", ", ".join(f"{n}={v:.2f}" for v, n in sorted(stalls, reverse=True)[:7]))


if __name__ == "__main__":
    main(sys.argv[1])
