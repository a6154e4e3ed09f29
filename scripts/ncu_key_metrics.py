"""Print the key counters of an `ncu --page raw --csv` export.  Usage: python scripts/ncu_key_metrics.py raw.csv"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
h, u, v = rows[0], rows[1], rows[2]
want = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__inst_issued.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__warps_eligible.avg.per_cycle_active", "smsp__issue_active.avg.per_cycle_active"]
print("kernel:", v[h.index("Kernel Name")][:100])
for i, n in enumerate(h):
    if n in want:
        print(f"  {n:85s} {v[i]} {u[i]}")
st = [(float(v[i]), n) for i, n in enumerate(h) if "stalled" in n and n.endswith("per_issue_active.ratio") and "not_issued" not in n]
print("  stalls (warps per issue):", ", ".join(f"{n.split('issue_stalled_')[1].split('_per_issue')[0]} {x:.2f}" for x, n in sorted(st, reverse=True)[:8]))
