import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sngnn_b200 import functional as SF
res = {}
for n, f, c in ((1632803, 65, 32), (1632803, 32, 2), (408201, 65, 32), (169343, 128, 32), (1632803, 128, 128)):
    gh = torch.randn(n, SF.padded_channels(c), device="cuda"); x = torch.randn(n, f, device="cuda")
    def timed(fn, steps=10, warm=3):
        for _ in range(warm): fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(steps): fn()
        b.record(); torch.cuda.synchronize()
        return round(a.elapsed_time(b) / steps, 4)
    res[f"n{n}_f{f}_c{c}"] = {"sng_ms": timed(lambda: SF.lin_bwd(gh, x, c)), "torch_ms": timed(lambda: (gh[:, :c].t() @ x, gh[:, :c].sum(0))),
                              "ideal_ms": round((n * (f + SF.padded_channels(c)) * 4) / 6.499e9 * 1e-3 * 1e3, 4)}
print(json.dumps(res))
