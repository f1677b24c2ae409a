"""Synthetic graphs and features of the shapes BASELINE.json names.

Pure torch, device-agnostic and seed-deterministic per device type; shared by the tests, `bench.py`
and `__graft_entry__.smoke()`.  No dataset download: only the *shapes* of the reference's datasets
matter here (R: datasets/largescale_datasets.py is not even importable as shipped).
"""
import torch

SHAPES = {
    # name: (num_nodes, num_features, num_directed_edges, num_classes)
    "tiny": (64, 16, 320, 4),
    "small": (1000, 48, 8000, 5),
    "chameleon": (2277, 2325, 36101, 5),
    "arxiv-year": (169343, 128, 1166243, 5),
    "pokec": (1632803, 65, 30622564, 2),
    "snap-patents": (2923922, 269, 13975788, 5),
}


def _gen(seed, device):
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    return g


def make_features(num_nodes, num_feats, kind="clustered", seed=0, device="cpu", dup_frac=0.001, zero_rows=8,
                  zscore=False, num_centroids=1024):
    """kind: 'normal' iid N(0,1) | 'clustered' (centroids + noise, intra-cluster cos ~ 0.9) | 'binary' (1 % dense 0/1).
    Adds exact duplicate rows (tie tests) and all-zero rows (F.normalize eps path)."""
    g = _gen(seed, device)
    N, Fd = num_nodes, num_feats
    if kind == "normal":
        x = torch.randn(N, Fd, generator=g, device=device)
    elif kind == "clustered":
        K = min(num_centroids, max(2, N // 8))
        c = torch.randn(K, Fd, generator=g, device=device)
        a = torch.randint(0, K, (N,), generator=g, device=device)
        x = c[a]
        x = x + 0.3 * x.norm(dim=1, keepdim=True) / Fd ** 0.5 * torch.randn(N, Fd, generator=g, device=device)
    elif kind == "binary":
        x = (torch.rand(N, Fd, generator=g, device=device) < 0.01).float()
    else:
        raise ValueError(kind)
    if zscore:  # R: datasets/largescale_datasets.py:462-465 (pokec node features are standardised)
        x = (x - x.mean(0, keepdim=True)) / x.std(0, keepdim=True).clamp(min=1e-6)
    nd = int(N * dup_frac)
    if nd:
        p = torch.randperm(N, generator=g, device=device)
        x[p[:nd]] = x[p[nd:2 * nd]]
    if zero_rows and N > 4 * zero_rows:
        z = torch.randperm(N, generator=g, device=device)[:zero_rows]
        x[z] = 0
    return x.contiguous()


def make_graph(num_nodes, num_edges, seed=1, device="cpu", symmetric=False, hub_offset=100.0, alpha=0.75):
    """Directed edges with power-law in-degree: dst ~ p_i ∝ (rank_i + hub_offset)^-alpha over a random node
    permutation, src uniform; coalesced and sorted row-major like PyG `coalesce` (R: datasets/datasets.py:170).
    Node 0 is forced to have an out-edge (R: models/models.py:125 shifts sources by min(src))."""
    g = _gen(seed, device)
    N = num_nodes
    E = num_edges // 2 if symmetric else num_edges
    w = (torch.arange(N, device=device, dtype=torch.float64) + hub_offset) ** (-alpha)
    cdf = torch.cumsum(w, 0)
    cdf = cdf / cdf[-1]
    u = torch.rand(E, generator=g, device=device, dtype=torch.float64)
    rank = torch.searchsorted(cdf, u).clamp(max=N - 1)
    perm = torch.randperm(N, generator=g, device=device)
    dst = perm[rank]
    src = torch.randint(0, N, (E,), generator=g, device=device)
    src[0], dst[0] = 0, (1 if N > 1 else 0)
    if symmetric:
        src, dst = torch.cat([src, dst]), torch.cat([dst, src])
    key = torch.unique(src * N + dst)          # sorted => row-major coalesced
    return torch.stack([key // N, key % N]).contiguous()


def make_labels(num_nodes, num_classes, seed=2, device="cpu"):
    return torch.randint(0, num_classes, (num_nodes,), generator=_gen(seed, device), device=device)


class GraphData:
    """Minimal stand-in for a PyG `Data` object: the reference models only read `.x` and `.edge_index`
    (R: models/models.py:77)."""

    def __init__(self, x, edge_index, y=None):
        self.x, self.edge_index, self.y = x, edge_index, y

    def to(self, device):
        return GraphData(self.x.to(device), self.edge_index.to(device), None if self.y is None else self.y.to(device))


def make_dataset(name, device="cpu", feature_kind=None, symmetric=True):
    N, Fd, E, C = SHAPES[name]
    kind = feature_kind or ("binary" if name == "chameleon" else "clustered")
    x = make_features(N, Fd, kind, device=device, zscore=(name == "pokec"))
    ei = make_graph(N, E, device=device, symmetric=symmetric)
    return GraphData(x, ei, make_labels(N, C, device=device)), C
