"""K2 forward / backward timing on the pokec-shaped graph: python scripts/k2_time.py [C] [top_k]"""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sngnn_b200 import synth, graph as G, functional as SF, _C
C = int(sys.argv[1]) if len(sys.argv) > 1 else 32
k = int(sys.argv[2]) if len(sys.argv) > 2 else 10
N, Fd, E, _ = synth.SHAPES["pokec"]
dev = "cuda"
ei = synth.make_graph(N, E, seed=1, device=dev, symmetric=True)
g = G.prepare(ei, N, True, structural=True)
torch.manual_seed(0)
h = torch.randn(N, C, device=dev)
gg = torch.randn(N, C, device=dev)
def timed(f, reps=10):
    for _ in range(3): f()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): f()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps
sel = {}
def fwd(): sel["o"] = SF.EdgeTopkAgg.apply(h, g, k, 0.0)
ms_f = timed(fwd)
out, ss, sw, sc = sel["o"]
dval, dnrm, dh = torch.zeros_like(h), torch.zeros_like(h), torch.empty_like(h)
_, _, inv_norm = SF.rownorm(h, want_f32=False, want_inv=True)
def bwd():
    dval.zero_(); dnrm.zero_()
    _C.check(_C.lib().sng_edge_agg_bwd(_C.ptr(h), _C.ptr(inv_norm), _C.ptr(gg), N, N, 0, C, C, _C.ptr(g.rowptr_in), _C.ptr(g.col_in), k, _C.ptr(ss), _C.ptr(sw),
                                       _C.ptr(sc), _C.ptr(g.inv_deg), _C.ptr(dval), _C.ptr(dnrm), _C.ptr(dh), _C.stream()), "bwd")
ms_b = timed(bwd)
Ep, nsel = g.num_edges, int(sc.sum())
bf = Ep * (4 * C + 4) + N * (8 * C + 8) + 8 * N * k
bb = nsel * (12 * C + 16) + 5 * N * 4 * C
print(json.dumps(dict(C=C, k=k, fwd_ms=round(ms_f, 3), fwd_gbs=round(bf / ms_f / 1e6), bwd_ms=round(ms_b, 3), bwd_gbs=round(bb / ms_b / 1e6), nsel=nsel,
                      checksum=float(out.double().abs().sum()), selsum=int(ss.clamp(min=0).long().sum()))))
