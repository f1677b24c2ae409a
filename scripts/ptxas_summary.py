"""Registers / spills per kernel from the ptxas logs of the last build: python scripts/ptxas_summary.py [filter]"""
import re, subprocess, sys, glob, os
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
flt = sys.argv[1] if len(sys.argv) > 1 else ""
for f in sorted(glob.glob(os.path.join(root, "sngnn_b200/csrc/build/*.ptxas.log"))):
    log = open(f).read()
    for b in re.split(r"ptxas info\s+: Compiling entry function '", log)[1:]:
        name = b.split("'")[0]
        dem = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()
        short = re.sub(r"\(.*", "", dem).replace("void sng::", "")
        if flt not in short:
            continue
        m = re.search(r"Used (\d+) registers", b)
        sp = re.search(r"(\d+) bytes spill stores, (\d+) bytes spill loads", b)
        sm = re.search(r"(\d+) bytes smem", b)
        print(short.ljust(64), "regs", m.group(1) if m else "?", "spill", sp.groups() if sp else "-", "smem", sm.group(1) if sm else 0)
