"""A few deterministic-backward launches on the pokec-shaped graph for ncu (python scripts/k2b_one.py [fused])."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sngnn_b200 import synth, graph as G, functional as SF
dev = "cuda"
N, Fd, E, _ = synth.SHAPES["pokec"]
ei = synth.make_graph(N, E, seed=1, device=dev, symmetric=True)
g = G.prepare(ei, N, True)
h = torch.randn(N, 32, device=dev); gg = torch.randn(N, 32, device=dev)
fused = len(sys.argv) > 1 and sys.argv[1] == "fused"
fuse = (torch.randn(N, 32, device=dev), torch.randn(32, device=dev), torch.full((1,), 0.5, device=dev), None) if fused else None
out, ss, sw, sq, sc, inv, diff = SF._edge_fwd(h, g, 0, 10, 0.0, True, fuse, want_q=True)
for _ in range(3):
    SF.edge_bwd(h, inv, gg, g, 10, ss, sw, sq, sc, fuse[2] if fused else None, diff)
torch.cuda.synchronize()
print("ok")
