"""Row sharding across GPUs (one process per GPU, torch.distributed; NCCL over NVLink on the B200 box).

The path shards by rows:
  * similarity build: rank r owns query rows [lo_r, hi_r); the only exchanged data is x-hat (FP16 for the tensor
    cores + FP32 for the exact rescore), all-gathered once per build; outputs stay sharded.
  * aggregation forward: rank r owns target rows [lo_r, hi_r) of the CSR-by-target; h is all-gathered per layer.
The reference has no distributed code at all; nothing here replaces a reference file.

The collective plumbing below is backend-agnostic (works on CPU tensors with gloo, which is how tests/ cover it);
the compute callables default to the CUDA kernels and can be injected by the tests.
"""
import torch
import torch.distributed as dist


def world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_world_size(), dist.get_rank()
    return 1, 0


def rows_per_rank(n, world_size):
    return (n + world_size - 1) // world_size


def shard_bounds(n, world_size, rank):
    """Contiguous equal shards of ceil(n / world) rows; the last ranks may own fewer (possibly zero) rows."""
    r = rows_per_rank(n, world_size)
    return min(n, rank * r), min(n, (rank + 1) * r)


def all_gather_rows(local, n, out=None, pad=None):
    """All-gather row shards produced by `shard_bounds` into a [n, ld] tensor (every rank gets all rows).
    `local` holds this rank's rows; shards are padded to equal size for all_gather_into_tensor."""
    ws, rank = world()
    if ws == 1:
        return local
    r = rows_per_rank(n, ws)
    ld = local.size(1)
    if pad is None:
        pad = torch.zeros(r, ld, dtype=local.dtype, device=local.device)
    if out is None:
        out = torch.empty(ws * r, ld, dtype=local.dtype, device=local.device)
    pad[: local.size(0)].copy_(local)
    dist.all_gather_into_tensor(out, pad)
    return out[:n]


def build_knn_sharded(x_local, n, top_k, thr=-1.0, remove_self=True, normalize=None, build=None):
    """Similarity-kNN of this rank's query rows against all n nodes.

    x_local: this rank's rows [hi-lo, d] of the feature matrix (shard_bounds(n, world, rank)).
    Returns (idx, sim, cnt) for the local rows (global column ids), exactly the rows [lo, hi) of the unsharded build."""
    ws, rank = world()
    lo, hi = shard_bounds(n, ws, rank)
    if x_local.size(0) != hi - lo:
        raise ValueError(f"rank {rank} must hold rows [{lo},{hi}) but got {x_local.size(0)} rows")
    if normalize is None or build is None:
        from . import simknn
        normalize = normalize or simknn.normalize_operands
        build = build or simknn.build_knn_normalized
    d = x_local.size(1)
    xf, xh = normalize(x_local)                       # K0 on the local shard only
    xf_all = all_gather_rows(xf, n)                   # FP32 x-hat for the exact rescore
    xh_all = all_gather_rows(xh, n)                   # FP16 x-hat for the tensor cores
    if hi == lo:
        return None
    return build(xf_all, xh_all, d, top_k, thr, remove_self, lo, hi)


def edge_agg_forward_sharded(h_local, graph, top_k=None, thr=None, agg=None):
    """out_1 rows [lo, hi) from the all-gathered h (forward only; the sharded backward needs a reduce-scatter of
    dL/dh and is not implemented yet -- DESIGN.md §multi-GPU)."""
    ws, rank = world()
    n = graph.n
    lo, hi = shard_bounds(n, ws, rank)
    h_all = all_gather_rows(h_local, n)
    shard = graph.row_slice(lo, hi)
    if agg is None:
        from . import functional as SF
        agg = SF.edge_topk_agg_rows
    return agg(h_all, shard, lo, top_k, thr)
