"""Summarise an .ncu-rep (raw page CSV) into the handful of metrics DESIGN.md / bench.py cite."""
import csv
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "smsp__warps_eligible.avg.per_cycle_active", "smsp__warps_active.avg.per_cycle_active",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__cycles_elapsed.max", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
]


def main(path):
    # an .ncu-rep, or the CSV of its raw page (`ncu -i x.ncu-rep --page raw --csv`)
    if path.endswith(".csv"):
        out = open(path).read()
    else:
        out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    print("| kernel | " + " | ".join(k for k in KEYS if k in hdr) + " |")
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")]
        vals = ["%s %s" % (r[hdr.index(k)], units[hdr.index(k)]) for k in KEYS if k in hdr]
        print("| %s | %s |" % (name[:48], " | ".join(vals)))


if __name__ == "__main__":
    main(sys.argv[1])
