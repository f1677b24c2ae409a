"""One-time graph preparation for the edge-restricted path.

The reference re-does this work every forward in Python/PyG (R: models/models.py:117-127, :234-236, :323):
append N self loops, optionally drop every src==dst edge, and (for ++) build a COO adjacency with
sources shifted by min(src).  Here it is done once per (edge_index, flags), cached, and turned into the
int32 CSR arrays the kernels consume:

  * CSR by TARGET (`rowptr_in`, `col_in`): in-edges of every node in ORIGINAL EDGE-POSITION order -- the
    tie-break of the selection rule is "lowest position first" (SURVEY.md §3.3), so the order is part of
    the contract;  `inv_deg[i] = 1 / max(indeg(i), 1)` is the PyG aggr='mean' denominator.
  * CSR by SHIFTED SOURCE (`rowptr_out`, `col_out`): out-neighbours t of node (src - min src), for
    out_0 = A @ W^T (R: :124-130);  `col_in_shift = col_in - min(src)` serves its transpose in backward.

CUDA `edge_index`: one call of `sng_graph_prepare` (stable radix sort by target / source, include/sng.h).  CPU tensors
(the host-logic and gloo tests): the same construction in torch, with a stable argsort.
"""
import weakref

import torch

_CACHE = {}
_CACHE_MAX = 8


class PreparedGraph:
    __slots__ = ("n", "num_edges", "rowptr_in", "col_in", "inv_deg", "src_shift", "rowptr_out", "col_out",
                 "col_in_shift", "dst_sorted", "_shards")

    def row_slice(self, lo, hi):
        """Row-sharded view (targets [lo, hi)) for multi-GPU aggregation: rowptr rebased to 0.  Cached per (lo, hi)."""
        cache = getattr(self, "_shards", None)
        if cache is None:
            cache = self._shards = {}
        if (lo, hi) in cache:
            return cache[(lo, hi)]
        g = cache[(lo, hi)] = PreparedGraph()
        g.n = hi - lo
        b, e = int(self.rowptr_in[lo]), int(self.rowptr_in[hi])
        g.num_edges = e - b
        g.rowptr_in = (self.rowptr_in[lo:hi + 1] - b).contiguous()
        g.col_in = self.col_in[b:e].contiguous()
        g.inv_deg = self.inv_deg[lo:hi].contiguous()
        g.src_shift = self.src_shift
        g.col_in_shift = None if self.col_in_shift is None else self.col_in_shift[b:e].contiguous()
        g.rowptr_out = g.col_out = g.dst_sorted = None
        if self.rowptr_out is not None:                      # by-(shifted-)source CSR rows [lo, hi) for out_0 = A @ W^T
            bo, eo = int(self.rowptr_out[lo]), int(self.rowptr_out[hi])
            g.rowptr_out = (self.rowptr_out[lo:hi + 1] - bo).contiguous()
            g.col_out = self.col_out[bo:eo].contiguous()
        return g


def _csr_from_keys(keys, vals, n):
    """Stable sort of `vals` by `keys` (< n) -> (rowptr int32 [n+1], vals_sorted)."""
    perm = torch.argsort(keys, stable=True)
    counts = torch.bincount(keys, minlength=n)
    rowptr = torch.zeros(n + 1, dtype=torch.int64, device=keys.device)
    torch.cumsum(counts, 0, out=rowptr[1:])
    return rowptr.to(torch.int32), vals[perm], perm


def process_edges(edge_index, num_nodes, remove_self_loops):
    """R: models/models.py:117-120 -- loops appended at the END, then (optionally) every src==dst dropped."""
    loop = torch.arange(num_nodes, dtype=edge_index.dtype, device=edge_index.device)
    ei = torch.cat([edge_index, torch.stack([loop, loop])], dim=1)
    if remove_self_loops:
        ei = ei[:, ei[0] != ei[1]]
    return ei


def prepare(edge_index, num_nodes, remove_self_loops, structural=False, processed=False):
    """Build (or fetch from cache) the PreparedGraph of `edge_index` [2,E] int64.

    remove_self_loops: False for base SNConv (R: :323), the `is_remove_self_loops` flag otherwise.
    structural: also build the by-source CSR needed by SNConv_plus_plus.
    processed: `edge_index` already went through `process_edges` (used by tests)."""
    key = (edge_index.data_ptr(), tuple(edge_index.shape), edge_index._version, str(edge_index.device),
           int(num_nodes), bool(remove_self_loops), bool(processed))
    hit = _CACHE.get(key)
    if hit is not None and hit[0]() is edge_index and (not structural or hit[1].rowptr_out is not None):
        return hit[1]
    if edge_index.numel() and int(edge_index.max()) >= num_nodes:
        raise ValueError("edge_index refers to a node id >= num_nodes")
    if num_nodes >= 2 ** 31 or edge_index.size(1) + num_nodes >= 2 ** 31:
        raise ValueError("graph too large for int32 CSR")
    if edge_index.is_cuda and not processed:
        g = _prepare_cuda(edge_index, int(num_nodes), bool(remove_self_loops), bool(structural))
        if len(_CACHE) >= _CACHE_MAX:
            _CACHE.pop(next(iter(_CACHE)))
        _CACHE[key] = (weakref.ref(edge_index), g)
        return g
    ei = edge_index if processed else process_edges(edge_index, num_nodes, remove_self_loops)
    src, dst = ei[0], ei[1]
    g = PreparedGraph()
    g.n = int(num_nodes)
    g.num_edges = int(src.numel())
    g.rowptr_in, col_in, _ = _csr_from_keys(dst, src, g.n)
    g.col_in = col_in.to(torch.int32).contiguous()
    deg = (g.rowptr_in[1:] - g.rowptr_in[:-1]).to(torch.float32)
    g.inv_deg = (1.0 / deg.clamp(min=1)).contiguous()
    g.src_shift = 0
    g.rowptr_out = g.col_out = g.col_in_shift = g.dst_sorted = None
    if structural:
        g.src_shift = int(src.min()) if src.numel() else 0            # R: models/models.py:125
        g.rowptr_out, col_out, _ = _csr_from_keys(src - g.src_shift, dst, g.n)
        g.col_out = col_out.to(torch.int32).contiguous()
        g.col_in_shift = (g.col_in - g.src_shift).contiguous()
    if len(_CACHE) >= _CACHE_MAX:
        _CACHE.pop(next(iter(_CACHE)))
    _CACHE[key] = (weakref.ref(edge_index), g)
    return g


def _prepare_cuda(edge_index, n, remove_self_loops, structural):
    from . import _C
    lib = _C.lib()
    ei = edge_index.contiguous()
    if ei.dtype != torch.int64:
        ei = ei.long()
    e = ei.size(1)
    dev = ei.device
    cap = e + n
    g = PreparedGraph()
    g.n = n
    rowptr_in = torch.empty(n + 1, dtype=torch.int32, device=dev)
    col_in = torch.empty(cap, dtype=torch.int32, device=dev)
    g.inv_deg = torch.empty(n, dtype=torch.float32, device=dev)
    rowptr_out = torch.empty(n + 1, dtype=torch.int32, device=dev) if structural else None
    col_out = torch.empty(cap, dtype=torch.int32, device=dev) if structural else None
    col_in_shift = torch.empty(cap, dtype=torch.int32, device=dev) if structural else None
    info = torch.zeros(2, dtype=torch.int32, device=dev)
    wbytes = lib.sng_graph_prepare_workspace_bytes(e, n)
    if wbytes == 0:
        raise ValueError("graph too large for int32 CSR")
    ws = torch.empty(wbytes, dtype=torch.uint8, device=dev)
    _C.check(lib.sng_graph_prepare(_C.ptr(ei), e, n, int(remove_self_loops), int(structural), _C.ptr(rowptr_in), _C.ptr(col_in),
                                   _C.ptr(g.inv_deg), _C.ptr(rowptr_out), _C.ptr(col_out), _C.ptr(col_in_shift), _C.ptr(info),
                                   _C.ptr(ws), wbytes, _C.stream()), "sng_graph_prepare")
    kept, shift = (int(v) for v in info.tolist())               # one sync: the arrays are narrowed to the kept edges
    g.num_edges = kept
    g.rowptr_in, g.col_in = rowptr_in, col_in[:kept]
    g.src_shift = shift if structural else 0
    g.rowptr_out = rowptr_out
    g.col_out = col_out[:kept] if structural else None
    g.col_in_shift = col_in_shift[:kept] if structural else None
    g.dst_sorted = None
    return g


def clear_cache():
    _CACHE.clear()
