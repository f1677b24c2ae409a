"""Adjacency-as-features similarity metrics (R: SimGFAToolbox/sparse.py).

The reference column-normalises a scipy CSC matrix M (nodes are COLUMNS, sparse.py:13), forms M^T M and densifies it row by
row (:34,70,109,139).  Here nothing is densified in FP32:
  * the N x N cosine matrix (cosine_similarity_sparse, node_similarity_sparse -- their contract IS the dense matrix) is one
    tensor-core contraction of the 0/1 columns held in FP16, which is EXACT (FP32 accumulators = common-neighbour counts),
    scaled by 1 / (|a_i| |a_j|) in the epilogue (sng_gemm_nt_f16);
  * the edge metrics (linked / neighbourhood) merge the two sorted index lists of an edge's columns (sng_sparse_col_cos);
  * the class metric is the closed form <S_a, S_b> of the per-class sums of the normalised columns.
A size guard keeps the FP16 operand and the FP32 result of the dense-output functions within device memory."""
import numpy as np
import torch

from .. import _C
from . import dense as D

_MAX_DENSE_BYTES = 96 << 30


def edge_index_to_sparse_csc_tensor(x, edge_index):
    """R: SimGFAToolbox/utils.py:5-11."""
    from scipy import sparse as sp
    n = len(x)
    row, col = edge_index[0].cpu().numpy(), edge_index[1].cpu().numpy()
    return sp.csc_matrix((np.full(len(row), 1), (row, col)), shape=(n, n))


def _csc(mat):
    """(indptr, indices, data, inv_norm) on the device of a scipy sparse / dense matrix, columns = nodes, indices sorted and
    duplicate-free; inv_norm[j] = 1 / max(|column j|, 1e-300) (sklearn normalize leaves an all-zero column at zero)."""
    from scipy import sparse as sp
    m = sp.csc_matrix(mat).astype(np.float64)
    m.sum_duplicates()
    m.sort_indices()
    nrm = np.sqrt(np.asarray(m.multiply(m).sum(axis=0)).ravel())
    inv = np.where(nrm > 0, 1.0 / np.where(nrm > 0, nrm, 1.0), 0.0)
    dev = torch.device("cuda", torch.cuda.current_device())
    t = lambda a, dt: torch.from_numpy(np.ascontiguousarray(a)).to(dt).to(dev)
    return t(m.indptr, torch.int32), t(m.indices, torch.int32), t(m.data, torch.float32), t(inv, torch.float32), m.shape


def _dense_cosine(mat):
    indptr, indices, data, inv, (r, c) = _csc(mat)
    if 2 * r * c + 4 * c * c > _MAX_DENSE_BYTES:
        raise RuntimeError(f"cosine matrix of {c} nodes does not fit the dense-output guard ({(2 * r * c + 4 * c * c) >> 30} GiB)")
    rp = (r + 7) // 8 * 8
    a = torch.zeros(c, rp, dtype=torch.float16, device=data.device)          # row j = column j of M (exact: small integers)
    cols = torch.repeat_interleave(torch.arange(c, device=data.device), (indptr[1:] - indptr[:-1]).long())
    a[cols, indices.long()] = data.half()
    return D.gemm_nt(a, a, r, inv, inv)


def cosine_similarity_sparse(mat):
    """R: sparse.py:8-14 -- returned dense [N, N] (the reference returns a scipy matrix that it densifies afterwards)."""
    return _dense_cosine(mat).cpu()


def node_similarity_sparse(x):
    """R: sparse.py:17-41 -- every entry (diagonal included) as [N*N, 1], and their mean."""
    sim = _dense_cosine(x)
    return sim.reshape(-1, 1).cpu(), sim.mean().cpu()


def _edge_cos(mat, edge_index):
    indptr, indices, data, inv, _ = _csc(mat)
    ei = edge_index.to(data.device)
    a, b = ei[0].to(torch.int32).contiguous(), ei[1].to(torch.int32).contiguous()
    s = torch.empty(a.numel(), dtype=torch.float32, device=data.device)
    _C.call("sng_sparse_col_cos", data, _C.ptr(indptr), _C.ptr(indices), _C.ptr(data), _C.ptr(inv), _C.ptr(a), _C.ptr(b), a.numel(), _C.ptr(s))
    return s, ei


def linked_node_similarity_sparse(x, edge_index):
    """R: sparse.py:44-77 -- edges are walked in the given (source-sorted) order."""
    s, _ = _edge_cos(x, edge_index)
    return s.reshape(-1, 1).cpu(), s.mean().cpu()


def neighborhood_similarity_sparse(x, edge_index):
    """R: sparse.py:80-119 -- per-node mean over its out-edges, mean over ALL N nodes."""
    s, ei = _edge_cos(x, edge_index)
    n = x.shape[1]
    w = D._per_source_mean(s, ei[0], n)
    return w.reshape(-1, 1).cpu(), (w.sum() / n).cpu()


def class_similarity_sparse(x, y):
    """R: sparse.py:122-152 -- K x K mean cosine between classes = <S_a, S_b> / (n_a n_b), S_a = sum of the normalised columns of class a."""
    indptr, indices, data, inv, (r, c) = _csc(x)
    yd = y.to(data.device).long()
    k = len(torch.unique(yd))
    cols = torch.repeat_interleave(torch.arange(c, device=data.device), (indptr[1:] - indptr[:-1]).long())
    sums = torch.zeros(k * r, dtype=torch.float64, device=data.device)
    sums.index_add_(0, yd[cols] * r + indices.long(), data.double() * inv.double()[cols])
    sums = sums.reshape(k, r)
    cnt = torch.bincount(yd, minlength=k).double()
    return ((sums @ sums.t()) / (cnt[:, None] * cnt[None, :])).float().cpu()
