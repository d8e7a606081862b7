#!/usr/bin/env python
"""Summarise an exported `ncu --page raw --csv` (+ optional `--page source --csv`) capture: the metrics the roofline uses."""
import csv
import sys

WANT = ["gpu__time_duration.sum", "smsp__inst_executed.sum", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "launch__grid_size", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sector_hit_rate.pct",
        "l1tex__t_sector_hit_rate.pct", "smsp__warps_eligible.avg.per_cycle_active",
        "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum", "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum",
        "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__inst_executed_op_shared_ld.sum", "sm__cycles_active.avg", "smsp__cycles_active.avg"]


def main(raw, src=None):
    rows = list(csv.reader(open(raw)))
    hdr, units, vals = rows[0], rows[1], rows[-1]
    out = {}
    for i, h in enumerate(hdr):
        if h in WANT or "issue_stalled" in h and h.endswith("per_issue_active.ratio"):
            try:
                out[h] = float(vals[i].replace(",", ""))
            except ValueError:
                out[h] = vals[i]
            if h in WANT:
                print(f"{h:75s} {vals[i]:>16s} {units[i]}")
    st = sorted(((v, k) for k, v in out.items() if "issue_stalled" in k), reverse=True)
    print("stalls per issue:", ", ".join(f"{k.split('issue_stalled_')[1].split('_per')[0]}={v:.2f}" for v, k in st[:8]))
    d = out
    fl = 2 * d.get("smsp__sass_thread_inst_executed_op_dfma_pred_on.sum", 0) + d.get("smsp__sass_thread_inst_executed_op_dmul_pred_on.sum", 0) + d.get("smsp__sass_thread_inst_executed_op_dadd_pred_on.sum", 0)
    print(f"fp64 flop (2*DFMA + DMUL + DADD thread instr): {fl:.4g}")
    if src:
        rows = list(csv.reader(open(src)))
        h = rows[0]
        ci = {n: i for i, n in enumerate(h)}
        print("source columns:", [c for c in h][:40])


if __name__ == "__main__":
    main(*sys.argv[1:])
