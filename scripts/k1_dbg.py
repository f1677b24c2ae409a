import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sngnn_b200 import simknn, synth
n, d, ew = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
x = synth.make_features(n, d, "normal", seed=1, device="cuda")
t = time.time()
ci, cv, cm, xf, xh = simknn.stage1_candidates(x, 14, thr_lo=-2.0, remove_self=True, force_ew=ew, force_nsplit=1)
torch.cuda.synchronize()
print("stage1 ok", n, d, ew, round(time.time() - t, 3), "lists", ci.shape, "filled", int((ci >= 0).sum()), "cm max", float(cm.max()), flush=True)
t = time.time()
idx, sim, cnt, nfb = simknn.build_knn(x, 10, -1.0, True, return_fallback=True)
torch.cuda.synchronize()
print("build ok", round(time.time() - t, 3), "fallback", int(nfb[0]), "cnt min", int(cnt.min()), flush=True)
