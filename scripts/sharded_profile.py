"""Kernel-time table (torch.profiler, rank 0) of the row-sharded SNGNN++ training step at the pokec shape.
torchrun --nproc-per-node N scripts/sharded_profile.py"""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
from torch.profiler import profile, ProfilerActivity
from sngnn_b200 import synth, dist as D, functional as SF
import sngnn_b200.models as M

world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local); dev = f"cuda:{local}"
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device(dev))
N, Fd, E, C = synth.SHAPES["pokec"]
x = synth.make_features(N, Fd, "clustered", seed=0, device=dev, zscore=True)
ei = synth.make_graph(N, E, seed=1, device=dev, symmetric=True)
y = synth.make_labels(N, C, seed=2, device=dev)
lo, hi = D.shard_bounds(N, world, rank)
torch.manual_seed(2)
model = M.SNGNN_Plus_Plus(Fd, 32, C, N, 2, 10, 0.0, 0.5, 1, 0.0).to(dev)
opt = torch.optim.Adam(model.parameters(), lr=0.01, weight_decay=5e-4, fused=True)
data = synth.GraphData(x, ei)
x_loc, y_loc = x[lo:hi].contiguous(), y[lo:hi]


def train_step():
    model.train(); opt.zero_grad()
    out = model(data) if world == 1 else D.sharded_forward(model, x_loc, ei, N)
    loss = SF.nll_loss(out, y) if world == 1 else SF.nll_loss(out, y_loc) * ((hi - lo) / N)
    loss.backward()
    D.allreduce_grads(model.parameters())
    opt.step()


for _ in range(3): train_step()
torch.cuda.synchronize()
if world > 1: dist.barrier()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(5): train_step()
b.record(); torch.cuda.synchronize()
ms = a.elapsed_time(b) / 5
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(3): train_step()
    torch.cuda.synchronize()
if rank == 0:
    rows = []
    for e in prof.key_averages():
        t = getattr(e, "device_time_total", 0) or getattr(e, "cuda_time_total", 0)
        if t > 0 and e.device_type.name == "CUDA":
            rows.append((t / 3 / 1e3, e.count // 3, e.key[:90]))
    rows.sort(reverse=True)
    print(f"world {world}: train step {ms:.3f} ms; kernel time per step (ms, launches, name):")
    tot = 0
    for t, c, k in rows[:40]:
        print(f"  {t:8.3f} {c:4d}  {k}"); tot += t
    print(f"  sum of top 40: {tot:.3f} ms")
if world > 1:
    dist.barrier(); dist.destroy_process_group()
