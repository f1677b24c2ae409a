"""A handful of edge-forward launches on the pokec-shaped graph for ncu (python scripts/k2_one.py [fused])."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sngnn_b200 import synth, graph as G, functional as SF
dev = "cuda"
N, Fd, E, _ = synth.SHAPES["pokec"]
ei = synth.make_graph(N, E, seed=1, device=dev, symmetric=True)
g = G.prepare(ei, N, True)
h = torch.randn(N, 32, device=dev)
fuse = None
if len(sys.argv) > 1 and sys.argv[1] == "fused":
    fuse = (torch.randn(N, 32, device=dev), torch.randn(32, device=dev), torch.full((1,), 0.5, device=dev), None)
for _ in range(4):
    SF._edge_fwd(h, g, 0, 10, 0.0, False, fuse)
torch.cuda.synchronize()
print("ok")
