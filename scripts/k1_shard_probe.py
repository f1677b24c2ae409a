"""One rank's share of a sharded build (query rows [0, N/parts)) with / without forced seeding: python scripts/k1_shard_probe.py N d k parts"""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sngnn_b200 import simknn, synth, _C
N, d, k, parts = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
x = synth.make_features(N, d, "clustered", seed=0, device="cuda")
xf, xh = simknn.normalize_operands(x)
hi = (N + parts - 1) // parts
def timed():
    run = lambda: simknn.build_knn_normalized(xf, xh, d, k, 0.0, True, 0, hi, return_fallback=True)
    run(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); r = run(); b.record(); torch.cuda.synchronize()
    return round(a.elapsed_time(b), 3), r[3].tolist()
res = {"nq": hi, "plan": simknn.build_plan(hi, N, d, k), "default": timed()}
_C.lib().sng_set_debug_env(1)
for s in (8, 16):
    os.environ["SNG_KNN_SEED_S"] = str(s)
    res[f"seed_stride_{s}"] = (simknn.build_plan(hi, N, d, k), timed())
print(json.dumps(res))
