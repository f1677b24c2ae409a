"""GPU toolbox (sngnn_b200.toolbox) against the golden outputs of the reference's own dense.py / sparse.py."""
import numpy as np
import pytest
import torch

from conftest import load_golden

pytestmark = pytest.mark.gpu


def _close(a, b, rtol=1e-4, atol=1e-5):
    torch.testing.assert_close(torch.as_tensor(a).float().cpu(), torch.as_tensor(b).float().cpu(), rtol=rtol, atol=atol)


def test_dense_metrics_match_reference_golden():
    import sngnn_b200.toolbox as T
    g = load_golden("toolbox.pt")
    x, ei, y = g["x"], g["edge_index"].long(), g["y"]
    _close(T.node_similarity_dense_large_parted(g["x1200"])[1], g["node_large_parted"], rtol=1e-3)
    _close(T.class_similarity_dense_large(g["x1200"], g["y1200"]), g["class_large_1200"])
    _close(T.class_similarity_dense_large(x, y), g["class_large"])
    for name, fn, args in (("linked_large", T.linked_node_similarity_dense_large, (x, ei)),
                           ("nbr_large", T.neighborhood_similarity_dense_large, (x, ei)),
                           ("node_small", T.node_similarity_dense_small, (x,)),
                           ("linked_small", T.linked_node_similarity_dense_small, (x, ei)),
                           ("nbr_small", T.neighborhood_similarity_dense_small, (x, ei)),
                           ("class_small", T.class_similarity_dense_small, (x, y))):
        got, ref = fn(*args), g[name]
        assert got[0].shape == ref[0].shape, name
        _close(got[0], ref[0])
        _close(got[1], ref[1])
    from sngnn_b200.toolbox.dense import edge_similarity_weight
    _close(edge_similarity_weight(g["esw_x"], ei), g["esw"])
    # device inputs stay on the device
    out = T.linked_node_similarity_dense_small(x.cuda(), ei.cuda())
    assert out[0].is_cuda


def test_sparse_metrics_match_reference_golden():
    import sngnn_b200.toolbox as T
    g = load_golden("toolbox.pt")
    x, ei, y = g["x"], g["edge_index"].long(), g["y"]
    adj = T.edge_index_to_sparse_csc_tensor(x, ei)
    for name, fn, args in (("sp_node", T.node_similarity_sparse, (adj,)), ("sp_linked", T.linked_node_similarity_sparse, (adj, ei)),
                           ("sp_nbr", T.neighborhood_similarity_sparse, (adj, ei))):
        got, ref = fn(*args), g[name]
        assert got[0].shape == ref[0].shape, name
        _close(got[0], ref[0])
        _close(got[1], ref[1])
    _close(T.class_similarity_sparse(adj, y), g["sp_class"])
    assert sorted(T.__all__) == sorted(
        ['cosine_similarity_sparse', 'node_similarity_sparse', 'linked_node_similarity_sparse', 'class_similarity_sparse',
         'plot_class_similarity', 'plot_similarity_distribution', 'edge_index_to_sparse_csc_tensor', 'node_similarity_dense_small',
         'node_similarity_dense_large_parted', 'class_similarity_dense_small', 'class_similarity_dense_large',
         'linked_node_similarity_dense_large', 'linked_node_similarity_dense_small', 'neighborhood_similarity_dense_large',
         'neighborhood_similarity_dense_small'])


def test_class_sums_large_shape():
    """Closed-form all-pairs sum at a size where the N x N route is impossible, against a float64 torch reduction."""
    import sngnn_b200.toolbox.dense as D
    torch.manual_seed(0)
    n, d, k = 300000, 65, 7
    x = torch.randn(n, d, device="cuda")
    y = torch.randint(0, k, (n,), device="cuda")
    got = D.class_similarity_dense_large(x, y).double()
    xh = torch.nn.functional.normalize(x.double(), dim=-1)
    s = torch.zeros(k, d, dtype=torch.float64, device="cuda").index_add_(0, y, xh)
    cnt = torch.bincount(y, minlength=k).double()
    ref = (s @ s.t()) / (cnt[:, None] * cnt[None, :])
    torch.testing.assert_close(got, ref, rtol=1e-4, atol=1e-9)


def test_cosine_similarity_noeps_and_knn_to_csr_and_segment_mean():
    """R: utils/data_transform.py:83-86 (x / ||x|| without eps: an all-zero row gives NaN rows / columns), the CSR emit of the
    kNN lists (sng_knn_to_csr) and the scatter_mean replacement (sng_segment_mean) against torch."""
    from oracle import toolbox_ref
    from sngnn_b200 import simknn
    import sngnn_b200.toolbox.dense as D
    torch.manual_seed(1)
    x = torch.randn(300, 17)
    x[5] = 0
    got = D.cosine_similarity(x)
    ref = toolbox_ref.cosine_similarity_noeps(x)
    assert torch.equal(torch.isnan(got), torch.isnan(ref))
    m = ~torch.isnan(ref)
    torch.testing.assert_close(got[m], ref[m], rtol=1e-5, atol=1e-6)
    # kNN lists -> CSR on the device == the host construction
    xd = torch.randn(2000, 33, device="cuda")
    idx, sim, cnt = simknn.build_knn(xd, 7, 0.2, True)
    rp, col, val = simknn.knn_to_csr(idx, cnt, sim)
    rp_h, col_h, val_h = simknn.knn_to_csr(idx.cpu(), cnt.cpu(), sim.cpu())
    assert torch.equal(rp.cpu(), rp_h) and torch.equal(col.cpu(), col_h) and torch.equal(val.cpu(), val_h)
    # segmented mean
    s = torch.randn(50000, device="cuda")
    seg = torch.randint(0, 997, (50000,), device="cuda")
    out = D._per_source_mean(s, seg, 1000)
    tot = torch.zeros(1000, dtype=torch.float64, device="cuda").index_add_(0, seg, s.double())
    c = torch.bincount(seg, minlength=1000).clamp(min=1)
    torch.testing.assert_close(out.double(), tot / c, rtol=1e-6, atol=1e-7)


@pytest.mark.parametrize("n,d", [(130, 7), (1000, 65), (3001, 300), (2277, 2325)])
def test_tensor_core_allpairs_matches_fp64(n, d):
    """cosine_similarity_dense_small through sng_gemm_nt_f16 (tcgen05, FP16 split hi + lo) against the FP64 product: the split
    keeps the FP32-level accuracy the goldens were generated with (ragged tile edges and K tails included)."""
    import sngnn_b200.toolbox.dense as D
    torch.manual_seed(n + d)
    x = torch.randn(n, d, device="cuda") * torch.rand(n, 1, device="cuda").mul(3).exp()
    x[3] = 0
    got = D.cosine_similarity_dense_small(x).double()
    xh = torch.nn.functional.normalize(x.double(), dim=-1)
    ref = xh @ xh.t()
    # the tensor cores accumulate K' = 3 d products per entry with truncating FP32 adds: ~2.5e-9 per product on unit rows
    assert (got - ref).abs().max().item() < 2e-6 + 2.5e-9 * 3 * d
    # plain exact use: small integers in FP16, FP32 accumulation, epilogue scales
    a = torch.randint(0, 3, (n, d), device="cuda").half()
    ap = torch.zeros(n, (d + 7) // 8 * 8, dtype=torch.float16, device="cuda")
    ap[:, :d] = a
    rs = torch.rand(n, device="cuda")
    cnt = D.gemm_nt(ap, ap, d, rs, rs)
    ref = (a.double() @ a.double().t()) * rs.double()[:, None] * rs.double()[None, :]
    torch.testing.assert_close(cnt.double(), ref, rtol=1e-6, atol=1e-6)


def test_sparse_metrics_mid_size_vs_oracle():
    """Adjacency-as-features metrics at a size where parity is more than the N = 300 golden: tensor-core M^T M with exact
    0/1 operands, merged-list edge cosines, closed-form class sums -- against the scipy oracle."""
    import sngnn_b200.toolbox as T
    from oracle import toolbox_ref
    from sngnn_b200 import synth
    n = 3000
    ei = synth.make_graph(n, 40000, seed=9, symmetric=True, hub_offset=3.0)
    ei = torch.cat([ei, ei[:, :300]], 1)                                  # duplicate edges: entries of M larger than 1
    ei = ei[:, (ei[0] * n + ei[1]).argsort(stable=True)]
    y = synth.make_labels(n, 4, seed=2)
    xdummy = torch.zeros(n, 1)
    adj = T.edge_index_to_sparse_csc_tensor(xdummy, ei)
    for fn, ref_fn, args in ((T.node_similarity_sparse, toolbox_ref.node_similarity_sparse, (adj,)),
                             (T.linked_node_similarity_sparse, toolbox_ref.linked_node_similarity_sparse, (adj, ei)),
                             (T.neighborhood_similarity_sparse, toolbox_ref.neighborhood_similarity_sparse, (adj, ei))):
        got, ref = fn(*args), ref_fn(*args)
        assert got[0].shape == ref[0].shape
        _close(got[0], ref[0], rtol=1e-5, atol=1e-6)
        _close(got[1], ref[1], rtol=1e-5, atol=1e-7)
    _close(T.class_similarity_sparse(adj, y), toolbox_ref.class_similarity_sparse(adj, y), rtol=1e-5, atol=1e-7)
