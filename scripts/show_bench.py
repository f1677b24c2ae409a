"""Compact view of a bench.py JSON line: python scripts/show_bench.py file.json"""
import json, sys
d = json.load(open(sys.argv[1]))
print({k: d[k] for k in ("value", "ms_per_step", "n_gpus")}, "e2e", d["e2e"]["value"], d["e2e"].get("copy_ms_rank0"))
print("roofline frac", d["roofline"]["frac"], "kernel ms", d["roofline"]["kernel_ms"], "| parity", d["parity"], "| clocks", d.get("clocks"))
print("cpu", d.get("cpu_baseline"))
for k, v in d.get("configs", {}).items():
    if "error" in v:
        print(k, v)
    elif "gpairs_per_s" in v:
        print(k, round(v["ms_per_step"], 1), "ms", round(v["gpairs_per_s"]), "Gp/s frac", round(v["main_pass_frac_of_tensor_peak"], 3), v["parity"])
    else:
        print(k, {kk: v[kk] for kk in ("epoch_ms_gpu", "epoch_ms_cpu_oracle", "parity")})
for k in ("epoch_ms", "forward_ms", "train_step_ms", "graph_prep_ms", "parity_epoch"):
    print(k, d.get(k))
a = d.get("agg")
if a:
    print({k: (round(v, 3) if isinstance(v, float) else v) for k, v in a.items() if k not in ("kernels", "algorithmic_bytes", "backward")})
