"""Summaries of EVERY kernel captured in an .ncu-rep (one block per launch): python scripts/ncu_all.py rep > profiles/x.summary.txt"""
import csv, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
r = list(csv.reader(raw.splitlines()))
keys = ["gpu__time_duration.sum", "smsp__inst_executed.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__warps_eligible.avg.per_cycle_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor.sum"]
units = dict(zip(r[0], r[1]))
for row in r[2:]:
    d = dict(zip(r[0], row))
    print("=== " + d.get("Kernel Name", "?"))
    for k in keys:
        if k in d:
            print(f"  {k:74s} {d[k]} {units.get(k, '')}")
    print("  stall reasons (warps per issue-active cycle):")
    for k, v in sorted(d.items()):
        if k.startswith("smsp__average_warps_issue_stalled_") and k.endswith("_per_issue_active.ratio"):
            try:
                if float(v) > 0.1:
                    print("     ", k.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "").ljust(28), v)
            except ValueError:
                pass
