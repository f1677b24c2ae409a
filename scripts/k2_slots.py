"""Forward edge kernels at the pokec shape for several slot-pool sizes of the staged kernel (SNG_K2_SLOTS)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sngnn_b200 import synth, graph as G, functional as SF, _C
dev = "cuda"
C = int(sys.argv[1]) if len(sys.argv) > 1 else 32
N, Fd, E, _ = synth.SHAPES["pokec"]
ei = synth.make_graph(N, E, seed=1, device=dev, symmetric=True)
g = G.prepare(ei, N, True)
torch.manual_seed(0)
h = torch.randn(N, C, device=dev)
fuse = (torch.randn(N, C, device=dev), torch.randn(C, device=dev), torch.full((1,), 0.5, device=dev), None)


def timed(fn, steps=10, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(steps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / steps


_C.lib().sng_set_debug_env(1)
res, ref = {}, {}
for slots in (0, 33, 40, 48, 56, 66, 65, 72, 80, 96, 112, 130):
    if slots: os.environ["SNG_K2_SLOTS"] = str(slots)
    for key, f in (("infer", lambda: SF._edge_fwd(h, g, 0, 10, 0.0, False)),
                   ("train", lambda: SF._edge_fwd(h, g, 0, 10, 0.0, True, want_q=True)),
                   ("fused_infer", lambda: SF._edge_fwd(h, g, 0, 10, 0.0, False, fuse)),
                   ("fused_train", lambda: SF._edge_fwd(h, g, 0, 10, 0.0, True, fuse, want_q=True))):
        if slots and ((slots < 65) != (not key.startswith("fused"))):
            continue
        r = f()
        if key not in ref: ref[key] = [t.clone() if torch.is_tensor(t) else t for t in r]
        else:
            for a, b in zip(ref[key], r):
                if torch.is_tensor(a): assert torch.equal(a, b), (slots, key)
        res[f"slots{slots}_{key}_ms"] = round(timed(f), 4)
_C.lib().sng_set_debug_env(0)
print(json.dumps(res))
