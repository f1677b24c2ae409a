"""Third-party shims so the UNMODIFIED reference (`/root/reference/models/models.py`,
`SimGFAToolbox/dense.py`, `sparse.py`) can be imported and executed on CPU in this
container, where torch_geometric / torch_scatter / torch_sparse / icecream are absent.

TEST INFRASTRUCTURE ONLY (dev-time golden-vector generation; see `oracle/make_golden.py`).
Nothing under `sngnn_b200/` may import this file.

The semantics below are the published behaviour of the versions the reference pins
(requirements.txt:66-69: torch-scatter 2.0.9, torch-sparse 0.6.13, torch-geometric 2.0.4),
restated from memory -- their sources are not under /root/reference, hence "parity unpinned"
for the third-party part (tie-break of scatter_max in particular: CPU rule = first position).
"""
import inspect
import sys
import types

import torch


# --------------------------------------------------------------------------- torch_scatter
def scatter_max(src, index, dim=0, out=None, dim_size=None):
    """torch_scatter.scatter_max, 1-D use at R: models/models.py:147,252.
    Output length index.max()+1; empty group -> (0, src.size(0)); ties -> first position."""
    assert src.dim() == 1 and dim in (0, -1)
    E = src.size(0)
    n = int(index.max()) + 1 if dim_size is None else dim_size
    if E == 0:
        return src.new_zeros(n), index.new_full((n,), E)
    val = src.new_full((n,), float("-inf")).scatter_reduce(0, index, src, "amax", include_self=True)
    is_max = src == val[index]
    pos = torch.arange(E, device=src.device)
    cand = torch.where(is_max, pos, torch.full_like(pos, E))
    arg = index.new_full((n,), E).scatter_reduce(0, index, cand, "amin", include_self=True)
    val = torch.where(arg == E, torch.zeros_like(val), val)
    return val, arg


def scatter_add(src, index, dim=0, out=None, dim_size=None):
    return scatter(src, index, dim, out, dim_size, reduce="sum")


def scatter_mean(src, index, dim=0, out=None, dim_size=None):
    return scatter(src, index, dim, out, dim_size, reduce="mean")


def scatter(src, index, dim=0, out=None, dim_size=None, reduce="sum"):
    """torch_scatter.scatter for dim 0 (and -1 on 1-D); 'mean' = sum / clamp(count, 1)."""
    if dim < 0:
        dim = src.dim() + dim
    assert dim == 0
    n = (int(index.max()) + 1 if index.numel() else 0) if dim_size is None else dim_size
    shape = (n,) + tuple(src.shape[1:])
    res = src.new_zeros(shape).index_add(0, index, src)
    if reduce in ("sum", "add"):
        return res
    if reduce == "mean":
        cnt = src.new_zeros(n).index_add(0, index, torch.ones(index.numel(), dtype=src.dtype))
        cnt = cnt.clamp(min=1)
        return res / cnt.view((n,) + (1,) * (src.dim() - 1))
    raise NotImplementedError(reduce)


# --------------------------------------------------------------------------- torch_sparse
class SparseTensor:
    """Only what R: models/models.py:126-127 uses (row/col/sparse_sizes -> COO with value 1)."""

    def __init__(self, row=None, col=None, value=None, sparse_sizes=None, **kw):
        self.row, self.col, self.value, self.sizes = row, col, value, sparse_sizes

    def to_torch_sparse_coo_tensor(self):
        v = self.value if self.value is not None else torch.ones(self.row.numel())
        return torch.sparse_coo_tensor(torch.stack([self.row, self.col]), v, tuple(self.sizes))


# --------------------------------------------------------------------------- torch_geometric
def add_self_loops(edge_index, edge_attr=None, fill_value=None, num_nodes=None):
    n = int(edge_index.max()) + 1 if num_nodes is None else num_nodes
    loop = torch.arange(n, dtype=edge_index.dtype, device=edge_index.device)
    return torch.cat([edge_index, torch.stack([loop, loop])], dim=1), edge_attr


def remove_self_loops(edge_index, edge_attr=None):
    mask = edge_index[0] != edge_index[1]
    return edge_index[:, mask], (None if edge_attr is None else edge_attr[mask])


def sort_edge_index(edge_index, edge_attr=None, num_nodes=None):
    n = int(edge_index.max()) + 1 if num_nodes is None else num_nodes
    perm = (edge_index[0] * n + edge_index[1]).argsort(stable=True)
    return edge_index[:, perm]


def maybe_num_nodes(edge_index, num_nodes=None):
    return int(edge_index.max()) + 1 if num_nodes is None else num_nodes


def zeros(t):
    if t is not None:
        t.data.fill_(0)


class MessagePassing(torch.nn.Module):
    """flow=source_to_target: `_j` = edge_index[0], `_i` = edge_index[1]; aggregate over
    edge_index[1] with dim_size = number of nodes (PyG 2.0.4 MessagePassing.propagate)."""

    def __init__(self, aggr="add", **kw):
        super().__init__()
        self.aggr = aggr

    def propagate(self, edge_index, size=None, **kwargs):
        src, dst = edge_index[0], edge_index[1]
        n = None
        for v in kwargs.values():
            if torch.is_tensor(v):
                n = v.size(0)
                break
        args = {}
        for name in inspect.signature(self.message).parameters:
            if name.endswith("_i"):
                args[name] = kwargs[name[:-2]].index_select(0, dst)
            elif name.endswith("_j"):
                args[name] = kwargs[name[:-2]].index_select(0, src)
            elif name == "index":
                args[name] = dst
            elif name == "edge_index":
                args[name] = edge_index
            else:
                args[name] = kwargs[name]
        msg = self.message(**args)
        return scatter(msg, dst, dim=0, dim_size=n, reduce=self.aggr)


class _Stub:
    def __init__(self, *a, **k):
        raise NotImplementedError("name-only stub")


def _module(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    m.__path__ = []
    m.__getattr__ = lambda attr: _Stub  # any other `from X import name` gets a stub
    sys.modules[name] = m
    return m


def install():
    """Register the fake modules (idempotent)."""
    if "torch_scatter" in sys.modules and getattr(sys.modules["torch_scatter"], "_sng_shim", False):
        return
    _module("icecream", ic=lambda *a, **k: None)
    _module("torch_scatter", scatter_max=scatter_max, scatter_add=scatter_add, scatter=scatter,
            scatter_mean=scatter_mean, _sng_shim=True)
    _module("torch_sparse", SparseTensor=SparseTensor)
    _module("torch_geometric")
    _module("torch_geometric.utils", add_self_loops=add_self_loops, remove_self_loops=remove_self_loops,
            sort_edge_index=sort_edge_index)
    _module("torch_geometric.utils.num_nodes", maybe_num_nodes=maybe_num_nodes)
    _module("torch_geometric.nn", MessagePassing=MessagePassing)
    _module("torch_geometric.nn.inits", zeros=zeros)
    _module("torch_geometric.nn.dense")
    _module("torch_geometric.nn.dense.linear")
    _module("torch_geometric.nn.conv")
    _module("torch_geometric.nn.conv.gcn_conv")
    _module("torch_geometric.typing")
    _module("torch_geometric.data")
    _module("torch_geometric.io")
    _module("torch_geometric.datasets")
    _module("torch_geometric.transforms")
    for m in ("matplotlib", "matplotlib.pyplot", "seaborn", "prettytable", "ogb", "ogb.nodeproppred", "gdown"):
        _module(m)
