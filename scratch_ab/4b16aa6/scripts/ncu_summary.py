"""Summarise an .ncu-rep (read here, no GPU needed) into a small text file for profiles/."""
import csv
import io
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sass__inst_executed_local_loads",
        "sass__inst_executed_local_stores", "sm__cycles_elapsed.avg", "lts__t_bytes.sum",
        "smsp__sass_inst_executed_op_global_ld.sum", "smsp__sass_inst_executed_op_global_red.sum"]


def main(rep, out):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    with open(out, "w") as f:
        f.write(f"# ncu summary of {rep}\n# (ncu --set full --clock-control none; per-launch, cold cache, serialised)\n")
        for r in rows[2:]:
            name = r[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?"
            f.write(f"\n## kernel: {name}\n")
            for h, u, v in zip(hdr, units, r):
                if h in KEYS or ("issue_stalled" in h and h.endswith("per_issue_active.ratio") and float(v or 0) > 0.1):
                    f.write(f"{h} [{u}] = {v}\n")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
