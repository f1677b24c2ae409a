"""Whole-build time at one shape for several seed strides (SNG_KNN_SEED_S): python scripts/k1_seed_probe.py N d k strides..."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sngnn_b200 import simknn, synth, _C
N, d, k = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
x = synth.make_features(N, d, "clustered", seed=0, device="cuda", zscore=(d == 65))
xf, xh = simknn.normalize_operands(x)
def timed(reps=3):
    run = lambda: simknn.build_knn_normalized(xf, xh, d, k, 0.0, True, return_fallback=True)
    run(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): r = run()
    b.record(); torch.cuda.synchronize()
    return round(a.elapsed_time(b) / reps, 2), r[3].tolist()
res = {"default": (simknn.build_plan(N, N, d, k)["seed_stride"], timed())}
_C.lib().sng_set_debug_env(1)
for s in sys.argv[4:]:
    os.environ["SNG_KNN_SEED_S"] = s
    p = simknn.build_plan(N, N, d, k)
    res[f"stride_{s}"] = ((p["seed_stride"], p["seed_q"]), timed())
print(json.dumps(res))
