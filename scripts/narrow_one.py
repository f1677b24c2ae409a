import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sngnn_b200 import synth, graph as G, functional as SF
dev = "cuda"; C = 4
N, Fd, E, _ = synth.SHAPES["pokec"]
ei = synth.make_graph(N, E, seed=1, device=dev, symmetric=True)
g = G.prepare(ei, N, True)
torch.manual_seed(0)
h = torch.randn(N, C, device=dev); gg = torch.randn(N, C, device=dev)
fuse = (torch.randn(N, C, device=dev), torch.randn(C, device=dev), torch.full((1,), 0.5, device=dev), None)
for _ in range(3):
    out, ss, sw, sq, sc, inv, diff = SF._edge_fwd(h, g, 0, 10, 0.0, True, fuse, want_q=True)
    SF.edge_bwd(h, inv, gg, g, 10, ss, sw, sq, sc, fuse[2], diff)
torch.cuda.synchronize(); print("ok")
