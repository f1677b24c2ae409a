"""Probe: torch symmetric memory / NVLS multicast on this box; multimem all-gather vs NCCL all-gather of h [N, 32] FP32."""
import sys, os, json
import torch, torch.distributed as dist
import torch.distributed._symmetric_memory as symm
world = int(os.environ["WORLD_SIZE"]); rank = int(os.environ["RANK"]); local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device(f"cuda:{local}")
dist.init_process_group("nccl", device_id=dev)
N, C = 1632803, 32
R = (N + world - 1) // world
group = dist.group.WORLD
buf = symm.empty((world * R, C), dtype=torch.float32, device=dev)
hdl = symm.rendezvous(buf, group)
if rank == 0:
    print("multicast support:", hdl.has_multicast_support, "mc ptr:", hex(hdl.multicast_ptr), "ptrs:", [hex(p) for p in hdl.buffer_ptrs][:4],
          "signal pad size", hdl.signal_pad_size, flush=True)
loc = torch.randn(R, C, device=dev) + rank
out = torch.empty(world * R, C, device=dev)


def timed(fn, steps=20, warm=5):
    for _ in range(warm): fn()
    torch.cuda.synchronize(); dist.barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(steps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / steps


t_nccl = timed(lambda: dist.all_gather_into_tensor(out, loc))
gname = group.group_name
try:
    t_mm = timed(lambda: torch.ops.symm_mem.multimem_all_gather_out(loc, gname, buf))
    ok = torch.equal(buf, out)
except Exception as e:
    t_mm, ok = None, repr(e)[:300]
t_bar = timed(lambda: hdl.barrier(channel=0))
red = torch.empty(R, C, device=dev)
full = torch.randn(world * R, C, device=dev)
t_rs = timed(lambda: dist.reduce_scatter_tensor(red, full))
if rank == 0:
    print(json.dumps({"world": world, "bytes_total": world * R * C * 4, "nccl_all_gather_ms": t_nccl, "multimem_all_gather_ms": t_mm, "equal": ok,
                      "symm_barrier_ms": t_bar, "nccl_reduce_scatter_ms": t_rs}), flush=True)
dist.barrier(); dist.destroy_process_group()
