"""K1 builds with wide features (streamed query block): time + plan.  python scripts/k1_wide.py"""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sngnn_b200 import synth, simknn
res = {}
for n, d, k in ((2277, 2325, 10), (100000, 768, 10), (100000, 1024, 10), (50000, 2325, 10)):
    x = synth.make_features(n, d, "clustered" if d != 2325 or n > 3000 else "binary", seed=1, device="cuda")
    xf, xh = simknn.normalize_operands(x)
    f = lambda: simknn.build_knn_normalized(xf, xh, d, k, 0.0, True, return_fallback=True)
    r = f(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(3): f()
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 3
    res[f"n{n}_d{d}"] = {"ms": round(ms, 3), "tflops": round(2.0 * n * n * d / ms / 1e9, 1), "plan": simknn.build_plan(n, n, d, k), "fallback": r[3].tolist()}
print(json.dumps(res))
