"""CPU oracle for the Sim-GFA toolbox similarity metrics.

TEST INFRASTRUCTURE ONLY (same import rule as `oracle/sn_ref.py`).

Vectorised restatement (torch / scipy on CPU) of what these reference functions RETURN,
including their quirks:
    R: SimGFAToolbox/dense.py:9-30     node_similarity_dense_large_parted   (formula precedence at :28)
    R: SimGFAToolbox/dense.py:33-62    linked_node_similarity_dense_large
    R: SimGFAToolbox/dense.py:65-101   neighborhood_similarity_dense_large  (divide by N incl. isolated, :96)
    R: SimGFAToolbox/dense.py:104-130  class_similarity_dense_large
    R: SimGFAToolbox/dense.py:138-179  *_small variants
    R: SimGFAToolbox/sparse.py:8-152   sparse (adjacency-as-features, COLUMN-normalised) variants
    R: utils/data_transform.py:83-91   cosine_similarity / edge_similarity_weight (no eps)
Parity pin: `tests/golden/toolbox_*.pt` hold the outputs of the reference's own dense.py /
sparse.py run through `oracle/shims.py`; `tests/test_oracle_golden.py` compares.
"""
import numpy as np
import torch
import torch.nn.functional as F


def _norm(x):
    return F.normalize(x, p=2.0, dim=-1)


def _sort_by_source(edge_index):
    n = int(edge_index.max()) + 1
    perm = (edge_index[0] * n + edge_index[1]).argsort(stable=True)
    return edge_index[:, perm]


def _edge_cos(n, src, dst):
    return (n[src] * n[dst]).sum(-1)


# ------------------------------------------------------------------------------------- dense
def node_similarity_dense_large_parted(x):
    n = _norm(x)
    N = n.size(0)
    total = _blocked_sum(n)
    return None, (total - N) / (N - 1) * N          # sic: R dense.py:28


def _blocked_sum(n, block=1000):
    tot = n.new_zeros(())
    for lo in range(0, n.size(0), block):
        tot = tot + (n[lo:lo + block] @ n.t()).sum()
    return tot


def linked_node_similarity_dense_large(x, edge_index):
    ei = _sort_by_source(edge_index)
    s = _edge_cos(_norm(x), ei[0], ei[1])
    return s.reshape(-1, 1), s.mean()


def neighborhood_similarity_dense_large(x, edge_index):
    ei = _sort_by_source(edge_index)
    n = _norm(x)
    N = n.size(0)
    s = _edge_cos(n, ei[0], ei[1])
    tot = torch.zeros(N).index_add(0, ei[0], s)
    deg = torch.zeros(N).index_add(0, ei[0], torch.ones_like(s))
    per_node = tot / deg.clamp(min=1)                 # isolated nodes contribute 0 / 1
    return per_node.reshape(-1, 1), per_node.sum() / N


def class_similarity_dense_large(x, y):
    n = _norm(x)
    K = len(torch.unique(y))
    onehot = F.one_hot(y.long(), K).to(n.dtype)       # [N,K]
    S = onehot.t() @ n                                # per-class sum of unit vectors [K,d]
    cnt = onehot.sum(0)
    return (S @ S.t()) / (cnt[:, None] * cnt[None, :])


def cosine_similarity_dense_small(x):
    n = _norm(x)
    return n @ n.t()


def node_similarity_dense_small(x):
    sim = cosine_similarity_dense_small(x)
    N = sim.size(0)
    off = sim[~torch.eye(N, dtype=torch.bool)]        # row-major off-diagonal, R dense.py:146-148
    return off, off.mean()


def linked_node_similarity_dense_small(x, edge_index):
    s = cosine_similarity_dense_small(x)[edge_index[0], edge_index[1]]
    return s.reshape(-1, 1), s.mean()


def neighborhood_similarity_dense_small(x, edge_index):
    s = _edge_cos(_norm(x), edge_index[0], edge_index[1])
    M = int(edge_index[0].max()) + 1                  # scatter_mean output length, R dense.py:163
    tot = torch.zeros(M).index_add(0, edge_index[0], s)
    deg = torch.zeros(M).index_add(0, edge_index[0], torch.ones_like(s))
    w = tot / deg.clamp(min=1)
    return w, w.mean()


def class_similarity_dense_small(x, y):
    m = class_similarity_dense_large(x, y)
    return m, m.mean()


# ------------------------------------------------------------------------------------ sparse
def cosine_similarity_sparse(mat):
    """R: sparse.py:8-14 -- columns are the nodes; L2-normalise columns, then M^T M (scipy CSC/CSR)."""
    import scipy.sparse as sp
    m = sp.csc_matrix(mat, dtype=np.float64)
    nrm = np.sqrt(np.asarray(m.multiply(m).sum(axis=0)).ravel())
    nrm[nrm == 0] = 1.0
    m = m @ sp.diags(1.0 / nrm)
    return (m.T @ m).tocsr()


def _dense_rows(sim):
    return torch.from_numpy(np.asarray(sim.todense(), dtype=np.float32))


def node_similarity_sparse(x):
    d = _dense_rows(cosine_similarity_sparse(x))
    N = d.size(0)
    return d.reshape(-1, 1), d.sum() / (N * N)


def linked_node_similarity_sparse(x, edge_index):
    d = _dense_rows(cosine_similarity_sparse(x))
    s = d[edge_index[0], edge_index[1]]               # caller passes source-sorted edges (R sparse.py:57-66)
    return s.reshape(-1, 1), s.mean()


def neighborhood_similarity_sparse(x, edge_index):
    ei = _sort_by_source(edge_index)
    d = _dense_rows(cosine_similarity_sparse(x))
    N = d.size(0)
    s = d[ei[0], ei[1]]
    tot = torch.zeros(N).index_add(0, ei[0], s)
    deg = torch.zeros(N).index_add(0, ei[0], torch.ones_like(s))
    per_node = tot / deg.clamp(min=1)
    return per_node.reshape(-1, 1), per_node.sum() / N


def class_similarity_sparse(x, y):
    d = _dense_rows(cosine_similarity_sparse(x))
    K = len(torch.unique(y))
    onehot = F.one_hot(y.long(), K).to(d.dtype)
    cnt = onehot.sum(0)
    return (onehot.t() @ d @ onehot) / (cnt[:, None] * cnt[None, :])


# -------------------------------------------------------------------------- utils/data_transform
def cosine_similarity_noeps(x):
    """R: utils/data_transform.py:83-86 -- x / ||x|| with no eps (zero rows -> NaN)."""
    x = x / torch.norm(x, dim=-1, keepdim=True)
    return x @ x.t()


def edge_similarity_weight(x, edge_index):
    """R: utils/data_transform.py:89-91."""
    return cosine_similarity_noeps(x)[edge_index[0].long(), edge_index[1].long()]
