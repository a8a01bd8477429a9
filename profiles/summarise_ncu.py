#!/usr/bin/env python
"""Turn `ncu --set full` reports (gpurun_out/*.ncu-rep, scratch) into the small CSV summaries committed here.

    python profiles/summarise_ncu.py gpurun_out/x.ncu-rep profiles/r2_ncu_x_summary.csv

One row per captured kernel launch, the metrics the design notes quote: duration, DRAM bytes, tensor-pipe / issue activity,
achieved occupancy, registers, L1 / L2 hit rates, shared-memory bank conflicts, instruction count and the stall reasons."""
import csv
import subprocess
import sys

METRICS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
           "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
           "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
           "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers",
           "launch__occupancy_limit_shared_mem", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
           "l1tex__throughput.avg.pct_of_peak_sustained_active", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
           "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
           "smsp__inst_executed.sum", "sm__inst_executed_pipe_tensor.sum"]


def main(rep, out):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    H, units = rows[0], rows[1]
    stalls = [h for h in H if "issue_stalled" in h and h.endswith("per_issue_active.ratio")]
    cols = ["Kernel Name"] + [m for m in METRICS if m in H] + stalls
    with open(out, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["kernel"] + [c + (" [" + units[H.index(c)] + "]" if units[H.index(c)] else "") for c in cols[1:]])
        for r in rows[2:]:
            w.writerow([r[H.index("Kernel Name")][:120]] + [r[H.index(c)] for c in cols[1:]])
    print(out, len(rows) - 2, "launches")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
