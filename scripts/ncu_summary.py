"""Key metrics + instruction mix + hottest SASS lines of the first kernel in an .ncu-rep: python scripts/ncu_summary.py rep [rows_for_per_row_counts]"""
import csv, subprocess, sys
from collections import defaultdict
rep = sys.argv[1]
nrows = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
r = list(csv.reader(raw.splitlines()))
d = dict(zip(r[0], r[2]))
keys = ["Kernel Name", "gpu__time_duration.sum", "smsp__inst_executed.sum", "launch__registers_per_thread", "launch__grid_size",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__warps_eligible.avg.per_cycle_active", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sector_hit_rate.pct",
        "l1tex__t_sector_hit_rate.pct", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.sum"]
for k in keys:
    print(k.ljust(72), d.get(k))
print("--- stall reasons (warps per issue-active cycle)")
for k, v in sorted(d.items()):
    if k.startswith("smsp__average_warps_issue_stalled_") and k.endswith("_per_issue_active.ratio"):
        try:
            if float(v) > 0.05:
                print("  ", k.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "").ljust(30), v)
        except ValueError:
            pass
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
hdr, out, nk = None, [], 0
for x in rows:
    if x and x[0] == "Kernel Name":
        nk += 1
    if x and x[0] == "Address":
        hdr = x
        continue
    if nk == 1 and hdr and len(x) > 5 and x[0].startswith("0x"):
        out.append(dict(zip(hdr, x)))
tot = sum(int(o["Instructions Executed"]) for o in out)
samp = sum(int(o["# Samples"]) for o in out) or 1
print(f"--- {len(out)} SASS lines, {tot} warp instructions ({tot / nrows:.1f} per row)")
agg, st = defaultdict(float), defaultdict(float)
for o in out:
    t = o["Source"].split()
    op = (t[1] if t[0].startswith("@") else t[0]).split(".")[0]
    agg[op] += int(o["Instructions Executed"]) / nrows
    st[op] += 100 * int(o["# Samples"]) / samp
for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:22]:
    print("  ", k.ljust(10), f"{v:10.1f}", f"{st[k]:5.1f}% samples")
print("--- hottest lines")
for o in sorted(out, key=lambda o: -int(o["# Samples"]))[:18]:
    print("  ", o["# Samples"].rjust(6), f"{int(o['Instructions Executed']) / nrows:8.2f}", o["Source"].strip()[:100])
