"""One full similarity-kNN build on the bench's pokec-shaped features (for ncu): seed pass + main pass + rescore + fallback."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sngnn_b200 import simknn, synth
N, Fd, E, C = synth.SHAPES[sys.argv[1] if len(sys.argv) > 1 else "pokec"]
x = synth.make_features(N, Fd, "clustered", seed=0, device="cuda", zscore=True)
idx, sim, cnt, nfb = simknn.build_knn(x, 10, 0.0, True, return_fallback=True)
torch.cuda.synchronize()
print("ok", int(nfb[0]), simknn.build_plan(N, N, Fd, 10))
