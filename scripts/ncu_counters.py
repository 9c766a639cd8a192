"""Turn the CSV of the ncu counter pass over scripts/probe_counters.py into
profiles/kernel_counters.json: FP64 lane-instructions and DRAM bytes per DoF of the cell kernel,
per degree, plain and with the in-loop vector updates.

  python scripts/ncu_counters.py gpurun_out/counters.csv gpurun_out/counters.log profiles/kernel_counters.json
"""
import csv
import json
import re
import sys


def main(csv_path, log_path, out_path):
    info = None
    for line in open(log_path):
        if line.startswith("COUNTER_INFO "):
            info = {int(k): v for k, v in json.loads(line[len("COUNTER_INFO "):]).items()}
    assert info, "COUNTER_INFO line missing"
    rows = [r for r in csv.reader(l for l in open(csv_path) if l.startswith('"'))]
    hdr = rows[0]
    ix = {h: i for i, h in enumerate(hdr)}
    per = {}                                    # (degree, fused) -> list of launches -> metric -> value
    for r in rows[1:]:
        m = re.search(r"cell_kernel<(?:\(int\))?(\d+), (?:\(int\))?\d+, (?:\(bool\))?(\d)(?:, (?:\(bool\))?\d)?>",
                      r[ix["Kernel Name"]])
        if not m:
            continue
        key = (int(m.group(1)), int(m.group(2)))
        lid = int(r[ix["ID"]])
        per.setdefault(key, {}).setdefault(lid, {})[r[ix["Metric Name"]]] = float(r[ix["Metric Value"]].replace(",", ""))
    out = {"source": "ncu --metrics pass over scripts/probe_counters.py (B200, --clock-control none)"}
    for (p, fused), launches in sorted(per.items()):
        last = launches[sorted(launches)[-1]]   # the launch after the warm-up one
        n = info[p]["n_dofs"]
        fp64 = last.get("smsp__thread_inst_executed_pipe_fp64_pred_on.sum", 0.0)
        dram = last.get("dram__bytes_read.sum", 0.0) + last.get("dram__bytes_write.sum", 0.0)
        out[f"q{p}_{'fused' if fused else 'plain'}"] = {
            "fp64_lane_instr_per_dof": fp64 / n, "dram_bytes_per_dof": dram / n,
            "dfma_per_dof": last.get("smsp__sass_thread_inst_executed_op_dfma_pred_on.sum", 0.0) / n,
            "dadd_per_dof": last.get("smsp__sass_thread_inst_executed_op_dadd_pred_on.sum", 0.0) / n,
            "dmul_per_dof": last.get("smsp__sass_thread_inst_executed_op_dmul_pred_on.sum", 0.0) / n,
            "fp64_pipe_active_pct": last.get("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed"),
            "kernel_us": last.get("gpu__time_duration.sum", 0.0) / 1e3,
            "n_dofs": n, "s": info[p]["s"], "source": "profiles/kernel_counters.json (ncu, s=%d)" % info[p]["s"]}
    json.dump(out, open(out_path, "w"), indent=1)
    for k, v in out.items():
        if isinstance(v, dict):
            print("%-10s fp64 lane-instr/DoF %7.1f  DRAM B/DoF %6.1f  fp64 pipe %s %%  %8.1f us" % (
                k, v["fp64_lane_instr_per_dof"], v["dram_bytes_per_dof"], v["fp64_pipe_active_pct"], v["kernel_us"]))


if __name__ == "__main__":
    main(*sys.argv[1:4])
