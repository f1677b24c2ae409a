"""K2 forward + backward and the K4 fusion once on the pokec-shaped graph (for ncu)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sngnn_b200 import synth, graph as G, functional as SF
N, Fd, E, C = synth.SHAPES[sys.argv[1] if len(sys.argv) > 1 else "pokec"]
dev = "cuda"
ei = synth.make_graph(N, E, seed=1, device=dev, symmetric=True)
g = G.prepare(ei, N, True, structural=True)
h = torch.randn(N, 32, device=dev, requires_grad=True)
w = torch.randn(32, N, device=dev, requires_grad=True)
bw = torch.zeros(32, device=dev, requires_grad=True)
beta = torch.full((1,), 0.5, device=dev, requires_grad=True)
for _ in range(2):
    out, _, _, _ = SF.EdgeTopkAgg.apply(h, g, 10, 0.0)
    o2 = SF.PPFuse.apply(out, w, bw, beta, None, g)
    o2.sum().backward()
torch.cuda.synchronize()
print("ok", g.num_edges)
