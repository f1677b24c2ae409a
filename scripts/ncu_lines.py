"""Per-CUDA-source-line warp-instruction counts of one kernel: joins the per-SASS counts of an .ncu-rep with nvdisasm -g
line info of the built library.  python scripts/ncu_lines.py rep mangled_name rows [min_per_row]"""
import csv, os, pickle, re, subprocess, sys, tempfile
from collections import defaultdict
rep, mangled, nrows = sys.argv[1], sys.argv[2], float(sys.argv[3])
thr = float(sys.argv[4]) if len(sys.argv) > 4 else 2.0
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.join(root, "sngnn_b200/lib/libsng.so")], cwd=tmp, capture_output=True)
cubin = [f for f in os.listdir(tmp) if f.startswith("sng_edge")][0] if "edge" in mangled else [f for f in os.listdir(tmp) if f.startswith("sng_simknn")][0]
sass = subprocess.run(["nvdisasm", "-g", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.split("\n")
start = next(i for i, l in enumerate(sass) if l.startswith(".text." + mangled + ":"))
seq, cur = [], None
for l in sass[start + 1:]:
    if l.startswith(".text.") or l.startswith("\t.section"):
        break
    m = re.search(r'//## File "[^"]*", line (\d+)', l)
    if m:
        cur = int(m.group(1)); continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m:
        seq.append(cur)
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
hdr, out, nk = None, [], 0
for x in rows:
    if x and x[0] == "Kernel Name":
        nk += 1
    if x and x[0] == "Address":
        hdr = x; continue
    if nk == 1 and hdr and len(x) > 5 and x[0].startswith("0x"):
        out.append(dict(zip(hdr, x)))
assert len(out) == len(seq), (len(out), len(seq))
cnt, smp = defaultdict(float), defaultdict(int)
for o, ln in zip(out, seq):
    cnt[ln] += int(o["Instructions Executed"]) / nrows
    smp[ln] += int(o["# Samples"])
text = open(os.path.join(root, "sngnn_b200/csrc/" + ("sng_edge.cu" if "edge" in mangled else "sng_simknn.cu"))).read().split("\n")
tot = sum(smp.values()) or 1
for ln in sorted(cnt, key=lambda k: (k is None, k)):
    if cnt[ln] >= thr:
        print(f"{ln!s:>5} {cnt[ln]:7.1f} {100 * smp[ln] / tot:5.1f}%  {text[ln - 1].strip()[:110] if ln else ''}")
print("total", sum(cnt.values()))
