"""Summarise an .ncu-rep (raw page) into a small text file for profiles/."""
import csv
import subprocess
import sys

rep, out = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, data = rows[0], rows[1], rows[2:]
keys = [
    "Kernel Name", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "gpu__time_duration.sum", "sm__cycles_active.avg", "smsp__cycles_active.avg",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__warps_active.avg.per_cycle_active", "smsp__warps_eligible.avg.per_cycle_active", "smsp__inst_executed.sum",
]
with open(out, "w") as f:
    f.write(f"# summary of {rep} (ncu --set full --clock-control none; per-launch values)\n")
    for r in data:
        rec = dict(zip(hdr, r))
        for k in keys:
            if k in rec:
                u = units[hdr.index(k)]
                f.write(f"{k:85s} {rec[k]} {u}\n")
        stalls = {h: float(v) for h, v in rec.items() if "issue_stalled" in h and h.endswith("per_issue_active.ratio") and v}
        for h, v in sorted(stalls.items(), key=lambda kv: -kv[1])[:8]:
            f.write(f"{h:85s} {v:.3f}\n")
        f.write("\n")
print(open(out).read())
