"""Adjacency-as-features similarity metrics (R: SimGFAToolbox/sparse.py).

The reference column-normalises a scipy CSC matrix M (nodes are COLUMNS, sparse.py:13) and forms M^T M, then densifies
it row by row (:34,70,109,139).  Here column j of M becomes the dense feature row of node j and the dense-feature kernels
are reused, so the values are the same cosines.  Everything these functions return or walk over is N x N in the reference
as well; a size guard keeps the densified features within device memory (the exact 0/1 tensor-core path for Penn94-scale
graphs is the "next" item of SURVEY.md §8(f))."""
import numpy as np
import torch

from . import dense as D

_MAX_DENSE_BYTES = 32 << 30


def edge_index_to_sparse_csc_tensor(x, edge_index):
    """R: SimGFAToolbox/utils.py:5-11."""
    from scipy import sparse as sp
    n = len(x)
    row, col = edge_index[0].cpu().numpy(), edge_index[1].cpu().numpy()
    return sp.csc_matrix((np.full(len(row), 1), (row, col)), shape=(n, n))


def _node_features(mat):
    """Dense [num_cols, num_rows] float32 tensor whose row j is column j of `mat` (scipy sparse or dense)."""
    if hasattr(mat, "tocoo"):
        coo = mat.tocoo()
        r, c = mat.shape
        if 4 * r * c > _MAX_DENSE_BYTES:
            raise RuntimeError(f"adjacency of shape {mat.shape} is too large to densify ({4 * r * c >> 30} GiB)")
        x = torch.zeros(c, r, dtype=torch.float32, device="cuda")
        idx = (torch.from_numpy(coo.col.astype(np.int64)).cuda(), torch.from_numpy(coo.row.astype(np.int64)).cuda())
        x.index_put_(idx, torch.from_numpy(coo.data.astype(np.float32)).cuda(), accumulate=True)
        return x
    return torch.as_tensor(mat, dtype=torch.float32).t().contiguous().cuda()


def cosine_similarity_sparse(mat):
    """R: sparse.py:8-14 -- returned dense [N, N] (the reference returns a scipy matrix that it densifies afterwards)."""
    return D.cosine_similarity_dense_small(_node_features(mat)).cpu()


def node_similarity_sparse(x):
    """R: sparse.py:17-41 -- every entry (diagonal included) as [N*N, 1], and their mean."""
    sim = D.cosine_similarity_dense_small(_node_features(x))
    return sim.reshape(-1, 1).cpu(), sim.mean().cpu()


def linked_node_similarity_sparse(x, edge_index):
    """R: sparse.py:44-77 -- edges are walked in the given (source-sorted) order."""
    s, m = D.linked_node_similarity_dense_small(_node_features(x), edge_index.cuda())
    return s.cpu(), m.cpu()


def neighborhood_similarity_sparse(x, edge_index):
    """R: sparse.py:80-119."""
    w, m = D.neighborhood_similarity_dense_large(_node_features(x), edge_index.cuda())
    return w.cpu(), m.cpu()


def class_similarity_sparse(x, y):
    """R: sparse.py:122-152."""
    return D.class_similarity_dense_large(_node_features(x), y.cuda()).cpu()
