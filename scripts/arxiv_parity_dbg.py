"""Where does the arxiv-shape SNGNN_Plus parity gap come from?  Layer-by-layer comparison with the oracle."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn.functional as F
from sngnn_b200 import synth, graph as G, functional as SF
import sngnn_b200.models as M
from oracle import sn_ref
dev = "cuda"
N, Fd, E, C = synth.SHAPES["arxiv-year"]
x = synth.make_features(N, Fd, "clustered", seed=0)
ei = synth.make_graph(N, E, seed=1, symmetric=True)
torch.manual_seed(2)
model = M.SNGNN_Plus(Fd, 32, C, N, 2, 10, 0.0, 1, 0.0)
sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
model = model.to(dev).eval()
g = G.prepare(ei.to(dev), N, True)
pe = sn_ref.process_edges(ei, N, True)
xin = x
for l, conv in enumerate(model.lins):
    h = conv._hidden(xin.to(dev))
    out, (ss, sw, sc) = SF.edge_topk_agg(h.requires_grad_(True), g, 10, 0.0, return_selection=True)
    hc = F.linear(xin, sd[f"lins.{l}.lin.weight"], sd[f"lins.{l}.lin.bias"])
    print("layer", l, "h err", float((h[:, :hc.size(1)].detach().cpu() - hc).abs().max() / hc.abs().max()))
    ref = sn_ref.sn_aggregate(hc, pe, 10, 0.0)
    o = out[:, :hc.size(1)].detach().cpu()
    rel = (o - ref).abs().max(1).values / ref.abs().max()
    bad = (rel > 1e-5).nonzero().flatten()
    print("  rows with rel err > 1e-5:", bad.numel(), "max", float(rel.max()))
    # selection lists of the oracle in FP64 on the SAME h (the GPU's h) to separate lin noise from selection
    h64 = h.detach().cpu().double()[:, :hc.size(1)]
    n64 = F.normalize(h64, dim=-1, eps=1e-12)
    s64 = (n64[pe[1]] * n64[pe[0]]).sum(-1)
    rank = sn_ref.edge_rank(s64, pe[1])
    selm = (rank < 10) & (s64 >= 0.0)
    idx_ref = torch.full((N, 10), -1, dtype=torch.long)
    idx_ref[pe[1][selm], rank[selm]] = pe[0][selm]
    cnt_ref = torch.zeros(N, dtype=torch.long).index_add(0, pe[1][selm], torch.ones(int(selm.sum()), dtype=torch.long))
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
    from parity import compare_lists
    res = compare_lists(ss, sc, idx_ref, cnt_ref, lambda r, j: (n64[r] * n64[j]).sum(-1), 0.0)
    print("  lists vs FP64 oracle on the GPU's own h:", res)
    same = ((ss.cpu().long() == idx_ref) | ((torch.arange(10)[None] >= sc.cpu()[:, None]) & (torch.arange(10)[None] >= cnt_ref[:, None]))).all(1)
    print("  bad rows whose list differs:", int((~same[bad]).sum()), "of", bad.numel())
    # FP32 oracle on GPU's h
    ref2 = sn_ref.sn_aggregate(h.detach().cpu()[:, :hc.size(1)], pe, 10, 0.0)
    rel2 = (o - ref2).abs().max(1).values / ref2.abs().max()
    print("  vs FP32 oracle on GPU's own h: rows > 1e-5:", int((rel2 > 1e-5).sum()), "max", float(rel2.max()))
    # zero rows in h?
    print("  zero rows in h:", int((h64.abs().sum(1) == 0).sum()), " duplicate-direction stats: min |h|", float(h64.norm(dim=1).min()))
    xin = F.relu(ref) if l == 0 else ref
    if l == 1:
        rp = g.rowptr_in.cpu()
        for r in bad.tolist():
            b, e = int(rp[r]), int(rp[r + 1])
            src = g.col_in[b:e].cpu().long()
            s = s64[(pe[1] == r)]
            print("row", r, "deg", e - b, "same list", bool(same[r]), "cnt", int(sc[r]), int(cnt_ref[r]))
            print("   ours", o[r].tolist()); print("   ref ", ref2[r].tolist())
            print("   list ours", ss[r].tolist()); print("   list ref ", idx_ref[r].tolist())
            print("   sel_w", sw[r].tolist())
            print("   s64 of ours", [(float((n64[r] * n64[j]).sum())) for j in ss[r].tolist() if j >= 0])
