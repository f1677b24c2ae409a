"""Per-kernel device time of one similarity-kNN build (torch.profiler): python scripts/k1_launches.py N d k [features]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
from sngnn_b200 import simknn, synth
N, d, k = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
feat = sys.argv[4] if len(sys.argv) > 4 else "clustered"
x = synth.make_features(N, d, feat, seed=0, device="cuda")
xf, xh = simknn.normalize_operands(x)
run = lambda: simknn.build_knn_normalized(xf, xh, d, k, 0.0, True, return_fallback=True)
for _ in range(2): run()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(3): run()
    torch.cuda.synchronize()
rows = sorted(((e.device_time_total / 3 / 1e3, e.count // 3, e.key[:100]) for e in prof.key_averages() if e.device_time_total > 0), reverse=True)
print(f"N={N} d={d} k={k} {feat}: kernel ms per build (ms, launches, name); sum = {sum(r[0] for r in rows):.3f} ms")
for t, c, kk in rows[:16]:
    print(f"  {t:8.3f} {c:3d}  {kk}")
