"""Times one similarity-kNN build shape: python scripts/k1_shape.py N d k [features] -> ms, fallback / retry rows, plan."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sngnn_b200 import simknn, synth
N, d, k = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
feat = sys.argv[4] if len(sys.argv) > 4 else "clustered"
x = synth.make_features(N, d, feat, seed=0, device="cuda")
xf, xh = simknn.normalize_operands(x)
def run():
    return simknn.build_knn_normalized(xf, xh, d, k, 0.0, True, return_fallback=True)
run(); torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record(); r = run(); b.record(); torch.cuda.synchronize()
ms = a.elapsed_time(b)
print(json.dumps({"N": N, "d": d, "k": k, "features": feat, "ms": ms, "gpairs": N * N / ms / 1e6, "fallback_rows": int(r[3][0]), "retry_rows": int(r[3][1]),
                  "plan": simknn.build_plan(N, N, d, k)}))
