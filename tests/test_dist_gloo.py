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
