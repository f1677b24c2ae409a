"""One SNGNN++ training step + one inference forward on the pokec-shaped graph, for a launch list:
   ncu --metrics gpu__time_duration.sum --csv --log-file out.csv python scripts/epoch_launches.py"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn.functional as F
from sngnn_b200 import synth
import sngnn_b200.models as M
dev = "cuda"
N, Fd, E, C = synth.SHAPES["pokec"]
x = synth.make_features(N, Fd, "clustered", seed=0, device=dev, zscore=True)
ei = synth.make_graph(N, E, seed=1, device=dev, symmetric=True)
y = synth.make_labels(N, C, seed=2, device=dev)
torch.manual_seed(2)
model = M.SNGNN_Plus_Plus(Fd, 32, C, N, 2, 10, 0.0, 0.5, 1, 0.5).to(dev)
opt = torch.optim.Adam(model.parameters(), lr=0.01, weight_decay=5e-4)
data = synth.GraphData(x, ei)
for it in range(2):
    model.train(); opt.zero_grad()
    torch.cuda.nvtx.range_push("train_step")
    F.nll_loss(model(data), y).backward(); opt.step()
    torch.cuda.nvtx.range_pop()
model.eval()
with torch.no_grad():
    model(data)
torch.cuda.synchronize()
print("ok")
