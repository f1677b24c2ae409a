"""world_size-2 gloo tests (CPU) of the row-sharding host logic in sngnn_b200/dist.py: shard bounds, padded all-gather,
and that a sharded build / aggregation equals the unsharded one.  The compute callables are the CPU oracle here (the
CUDA kernels are covered by the -m gpu tests); what is under test is the partitioning and the collective plumbing."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, ws, port, n, ret):
    import sys
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=ws)
    from sngnn_b200 import dist as D, graph as G, synth
    from oracle import sn_ref
    torch.manual_seed(0)
    d, k, thr = 12, 5, 0.1
    x = synth.make_features(n, d, "clustered", seed=3)
    lo, hi = D.shard_bounds(n, ws, rank)

    # --- padded all-gather reproduces the full matrix on every rank
    full = D.all_gather_rows(x[lo:hi].clone(), n)
    assert torch.equal(full, x)
    work, full2, keep = D._gather_row_blocks_async(x[lo:hi].clone(), n)   # the started-early form (NCCL: async handle; gloo: done at once)
    if work is not None:
        work.wait()
    assert torch.equal(full2, x)

    # --- sharded kNN build == rows [lo,hi) of the unsharded build
    def normalize(xs):
        nx = sn_ref.rownorm(xs)
        return nx, nx.half()

    def build(xf_all, xh_all, dd, top_k, th, rs, qlo, qhi):
        assert xh_all.dtype == torch.float16 and xf_all.size(0) == n
        return sn_ref.simknn_allpairs(xf_all, top_k, th, rs, qlo, qhi)      # x-hat is already unit norm

    got = D.build_knn_sharded(x[lo:hi].clone(), n, k, thr, True, normalize=normalize, build=build)
    ref = sn_ref.simknn_allpairs(x, k, thr, True)
    if hi > lo:
        assert torch.equal(got[0], ref[0][lo:hi]) and torch.equal(got[2], ref[2][lo:hi])
        assert torch.allclose(got[1], ref[1][lo:hi], atol=1e-6)           # x-hat is re-normalised inside the oracle

    # --- sharded aggregation forward == rows [lo,hi) of the unsharded oracle
    ei = synth.make_graph(n, 6 * n, seed=5, hub_offset=2.0)
    g = G.prepare(ei, n, True)
    h = torch.randn(n, 8)

    def agg(h_all, shard, row_offset, top_k, th):
        pe = sn_ref.process_edges(ei, n, True)
        out = sn_ref.sn_aggregate(h_all, pe, top_k, th)
        # the shard's CSR must describe exactly the in-edges of rows [row_offset, row_offset + shard.n)
        for i in (0, shard.n // 2, shard.n - 1):
            if shard.n:
                want = pe[0][pe[1] == row_offset + i]
                have = shard.col_in[shard.rowptr_in[i]:shard.rowptr_in[i + 1]].long()
                assert torch.equal(want, have)
        return out[row_offset:row_offset + shard.n]

    part = D.edge_agg_forward_sharded(h[lo:hi].clone(), g, k, thr, agg=agg)
    whole = sn_ref.sn_aggregate(h, sn_ref.process_edges(ei, n, True), k, thr)
    assert torch.allclose(part, whole[lo:hi])
    ret[rank] = (lo, hi)
    dist.barrier()
    dist.destroy_process_group()


def _run(n, ws=2):
    ctx = mp.get_context("spawn")
    ret = ctx.Manager().dict()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, ws, port, n, ret)) for r in range(ws)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
        assert p.exitcode == 0, f"worker exited with {p.exitcode}"
    return dict(ret)


def test_row_sharding_world2_even_and_ragged():
    assert _run(200) == {0: (0, 100), 1: (100, 200)}
    assert _run(201) == {0: (0, 101), 1: (101, 201)}          # ragged last shard exercises the padding


def test_shard_bounds_cover_everything():
    from sngnn_b200 import dist as D
    for n in (1, 7, 8, 1632803):
        for ws in (1, 2, 4, 8):
            b = [D.shard_bounds(n, ws, r) for r in range(ws)]
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(ws - 1))


def _train_worker(rank, ws, port, kind, n, ret):
    """Sharded forward + backward of a replicated-parameter model == the unsharded oracle (rows and parameter gradients)."""
    import sys
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=ws)
    import torch.nn.functional as F
    from sngnn_b200 import dist as D, synth
    import sngnn_b200.models as M
    from oracle import sn_ref
    fd, hid, ncls, k, thr = 10, 8, 3, 3, 0.0
    x = synth.make_features(n, fd, "clustered", seed=1).double()
    ei = synth.make_graph(n, 5 * n, seed=2, hub_offset=2.0)
    y = synth.make_labels(n, ncls, seed=3)
    torch.manual_seed(7)                                         # same replicated parameters on every rank
    if kind == "SNGNN":
        model = M.SNGNN(fd, hid, ncls, 2)
        model.dropout = torch.nn.Dropout(0.0)
    elif kind == "SNGNN_Plus":
        model = M.SNGNN_Plus(fd, hid, ncls, n, 2, k, thr, 1, 0.0)
    else:
        model = M.SNGNN_Plus_Plus(fd, hid, ncls, n, 2, k, thr, 0.5, 1, 0.0)
    model = model.double().train()
    lo, hi = D.shard_bounds(n, ws, rank)
    rsl = False if kind == "SNGNN" else True
    pe = sn_ref.process_edges(ei, n, rsl)

    def agg(h_all, shard, row_offset, top_k, th):              # differentiable CPU oracle of K2 on the shard's rows
        return sn_ref.sn_aggregate(h_all, pe, top_k, th)[row_offset:row_offset + shard.n]

    def fuse(out1, w_w, w_b, beta, bias, shard, nn_, lo_):      # differentiable CPU oracle of K4 on the shard's rows
        out0 = sn_ref.structural_term(pe, w_w, w_b, nn_)[lo_:lo_ + out1.size(0)]
        out0 = F.pad(out0, (0, out1.size(1) - out0.size(1)))
        out = beta * out0 + (1 - beta) * out1
        return out if bias is None else out + F.pad(bias, (0, out1.size(1) - bias.numel()))

    logp = D.sharded_forward(model, x[lo:hi].clone(), ei, n, agg=agg, fuse=fuse)
    loss = F.nll_loss(logp, y[lo:hi], reduction="sum") / n     # local part of the full-batch mean
    loss.backward()
    D.allreduce_grads(model.parameters())

    sd = {kk: v.detach().clone().requires_grad_(v.is_floating_point()) for kk, v in model.state_dict().items()}
    ref = sn_ref.stack_forward(kind, sn_ref.params_from_state_dict(sd, 2), x, ei, top_k=k, thr=thr, remove_self_loops=rsl)
    F.nll_loss(ref, y).backward()
    assert torch.allclose(logp.detach(), ref.detach()[lo:hi], rtol=1e-9, atol=1e-11)
    for name, p in model.named_parameters():
        g_ref = sd[name].grad
        assert p.grad is not None and torch.allclose(p.grad, g_ref, rtol=1e-8, atol=1e-11), (kind, name, (p.grad - g_ref).abs().max())
    ret[rank] = True
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_training_step_world2_matches_unsharded_oracle():
    """The N>1 training path: all-gather of h with a reduce-scatter backward, shard-local aggregation and
    ++ fusion, all-reduce of the partial parameter gradients -- against the unsharded oracle, ragged shards included."""
    ctx = mp.get_context("spawn")
    for kind, n in (("SNGNN_Plus_Plus", 61), ("SNGNN_Plus", 40), ("SNGNN", 33)):
        ret = ctx.Manager().dict()
        port = _free_port()
        procs = [ctx.Process(target=_train_worker, args=(r, 2, port, kind, n, ret)) for r in range(2)]
        for p in procs:
            p.start()
        for p in procs:
            p.join(180)
            assert p.exitcode == 0, f"{kind}: worker exited with {p.exitcode}"
        assert dict(ret) == {0: True, 1: True}


def _toolbox_worker(rank, ws, port, n, ret):
    import sys
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=ws)
    import torch.nn.functional as F
    from sngnn_b200 import dist as D, synth
    from sngnn_b200.toolbox import sharded as TS
    from oracle import toolbox_ref
    d, k = 9, 4
    x = synth.make_features(n, d, "clustered", seed=3)
    y = synth.make_labels(n, k, seed=4)
    ei = synth.make_graph(n, 5 * n, seed=5, hub_offset=2.0)
    lo, hi = D.shard_bounds(n, ws, rank)

    def class_sums(xl, yl, kk):                                  # CPU stand-in of K0 + sng_class_sums_f64 on the local rows
        xh = F.normalize(xl.float(), dim=-1).double()
        yy = torch.zeros(xl.size(0), dtype=torch.long) if yl is None else yl
        s = torch.zeros(kk, xl.size(1), dtype=torch.float64).index_add_(0, yy, xh)
        return s, torch.bincount(yy, minlength=kk).double()

    _, m = TS.node_similarity_dense_large_parted_sharded(x[lo:hi], n, class_sums=class_sums)
    assert torch.allclose(m, toolbox_ref.node_similarity_dense_large_parted(x)[1], rtol=1e-4)
    cs = TS.class_similarity_dense_large_sharded(x[lo:hi], y[lo:hi], k, class_sums=class_sums)
    assert torch.allclose(cs, toolbox_ref.class_similarity_dense_large(x, y), rtol=1e-4, atol=1e-6)
    s, mean = TS.linked_node_similarity_dense_sharded(x[lo:hi], ei, n, edge_cos=lambda xh, a, b: (xh[a] * xh[b]).sum(-1))
    ref_s, ref_mean = toolbox_ref.linked_node_similarity_dense_small(x, ei)
    per = (ei.size(1) + ws - 1) // ws
    assert torch.allclose(s, ref_s.flatten()[rank * per:(rank + 1) * per], atol=1e-6) and torch.allclose(mean, ref_mean, atol=1e-6)
    ret[rank] = True
    dist.barrier()
    dist.destroy_process_group()


def test_toolbox_reductions_world2_match_unsharded_oracle():
    """Sim-GFA metrics from row shards (SURVEY.md §8(e) 'toolbox reductions'): class sums all-reduced, edge means all-reduced."""
    ctx = mp.get_context("spawn")
    ret = ctx.Manager().dict()
    port = _free_port()
    procs = [ctx.Process(target=_toolbox_worker, args=(r, 2, port, 57, ret)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0, f"worker exited with {p.exitcode}"
    assert dict(ret) == {0: True, 1: True}
