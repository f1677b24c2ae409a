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


def half_operand(xf_all, ldh):
    """FP16 copy [n, ldh] (zero padded) of normalised FP32 rows [n, ld32]: the K1 tensor-core operand, converted locally."""
    xh = torch.zeros(xf_all.size(0), ldh, dtype=torch.float16, device=xf_all.device)
    w = min(ldh, xf_all.size(1))
    xh[:, :w] = xf_all[:, :w]
    return xh


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
    xf_all = all_gather_rows(xf, n)                   # ONE collective: FP32 x-hat (needed for the exact rescore anyway)
    # the tensor-core operand is the FP16 rounding of the same values, so it is converted locally instead of gathered:
    # bit-identical to what the owning rank's K0 emitted, and it saves the second all-gather (261 of 705 MB at pokec scale)
    xh_all = half_operand(xf_all, xh.size(1))
    if hi == lo:
        return None
    return build(xf_all, xh_all, d, top_k, thr, remove_self, lo, hi)


def edge_agg_forward_sharded(h_local, graph, top_k=None, thr=None, agg=None):
    """out_1 rows [lo, hi) from the all-gathered h, forward only (no autograd; see `edge_agg_sharded` for training)."""
    ws, rank = world()
    n = graph.n
    lo, hi = shard_bounds(n, ws, rank)
    h_all = all_gather_rows(h_local, n)
    shard = graph.row_slice(lo, hi)
    if agg is None:
        from . import functional as SF
        agg = SF.edge_topk_agg_rows
    return agg(h_all, shard, lo, top_k, thr)


# ------------------------------------------------------------------------------------------ training: sharded forward + backward
def _reduce_scatter_rows(full, n):
    """Sum `full` [ws*R, ld] (R = rows_per_rank) over ranks and return this rank's rows.  NCCL: reduce_scatter_tensor;
    backends without it (gloo on CPU, used by the tests): all_reduce + slice."""
    ws, rank = world()
    r = rows_per_rank(n, ws)
    lo, hi = shard_bounds(n, ws, rank)
    if dist.get_backend() == "nccl":
        out = torch.empty(r, full.size(1), dtype=full.dtype, device=full.device)
        dist.reduce_scatter_tensor(out, full.contiguous(), op=dist.ReduceOp.SUM)
        return out[: hi - lo]
    dist.all_reduce(full, op=dist.ReduceOp.SUM)
    return full[lo:hi]


class AllGatherRows(torch.autograd.Function):
    """h_all [n, C] = concatenation of every rank's row shard; backward = reduce-scatter (sum) of dL/dh_all -- the one
    exchange of the sharded backward: a rank's targets gather from, and so send gradient to, any source row."""

    @staticmethod
    def forward(ctx, local, n):
        ctx.n = n
        return all_gather_rows(local.contiguous(), n)

    @staticmethod
    def backward(ctx, g_all):
        ws, _ = world()
        n = ctx.n
        if ws == 1:
            return g_all, None
        r = rows_per_rank(n, ws)
        full = g_all.new_zeros(ws * r, g_all.size(1))
        full[:n].copy_(g_all)
        return _reduce_scatter_rows(full, n).contiguous(), None


def edge_agg_sharded(h_local, graph, top_k=None, thr=None, agg=None):
    """Differentiable out_1 rows [lo, hi): all-gather h (autograd-aware), then K2 on the shard; its backward produces the
    shard's contribution to dL/dh of every node, which AllGatherRows.backward reduce-scatters.
    `agg(h_all, shard, row_offset, top_k, thr)` defaults to the CUDA kernels (functional.ShardedEdgeTopkAgg); the gloo tests
    inject a differentiable CPU oracle."""
    ws, rank = world()
    n = graph.n
    lo, hi = shard_bounds(n, ws, rank)
    h_all = AllGatherRows.apply(h_local, n)
    shard = graph.row_slice(lo, hi)
    if agg is None:
        from . import functional as SF
        agg = SF.ShardedEdgeTopkAgg.apply
    return agg(h_all, shard, lo, top_k, thr), shard


def pp_fuse_sharded(out1_local, w_weight, w_bias, beta, bias, shard, n, lo, fuse=None):
    """SNGNN++ fusion on a row shard: out rows [lo, hi) = beta * (A[lo:hi] @ W^T + b_w) + (1 - beta) * out_1 (+ bias), with
    the replicated parameter W^T [n, C] gathered over the shard's out-neighbours.  The parameter gradients this produces
    are PARTIAL (this rank's rows only), like every other parameter gradient of a sharded step: `allreduce_grads` sums them.
    `fuse` defaults to functional.PPFuse on the shard's by-source CSR (its backward needs g0 of ALL rows for dL/dW^T rows
    [lo, hi), so that variant all-gathers g0); the gloo tests inject a differentiable dense CPU version."""
    if fuse is not None:
        return fuse(out1_local, w_weight, w_bias, beta, bias, shard, n, lo)
    return _ShardedPPFuse.apply(out1_local, w_weight, w_bias, beta, bias, shard, n, lo)


def _gather_row_blocks(local, n):
    """[n, ld] from every rank's [hi - lo, ld] block (no autograd)."""
    return all_gather_rows(local.contiguous(), n)


class _ShardedPPFuse(torch.autograd.Function):
    """Unfused SNGNN++ epilogue on a row shard (graphs whose in-lists differ from their out-lists): sng_pp_fuse_fwd over the
    shard's by-source CSR.  dL/dW^T rows [lo, hi) need g0 of ALL rows (all-gather of g0); the row blocks of the gradient
    are then all-gathered, so every rank ends up with the COMPLETE dL/dw.weight (no all-reduce of a zero-padded matrix)."""

    @staticmethod
    def forward(ctx, out1, w_weight, w_bias, beta, bias, shard, n, lo):
        import torch.nn.functional as F
        from . import _C, functional as SF
        out1 = SF._check_h(out1)
        nl, cp = out1.shape
        c = w_weight.size(0)
        wt = SF._padded_wt(w_weight, cp)                                          # [n, Cp], replicated
        bw = F.pad(w_bias.detach(), (0, cp - c)).contiguous()
        bb = None if bias is None else F.pad(bias.detach(), (0, cp - c)).contiguous()
        out0, out = torch.empty_like(out1), torch.empty_like(out1)
        if nl:
            _C.call("sng_pp_fuse_fwd", out1, _C.ptr(wt), nl, cp, cp, _C.ptr(shard.rowptr_out), _C.ptr(shard.col_out), _C.ptr(bw),
                    _C.ptr(beta), _C.ptr(out1), _C.ptr(bb), _C.ptr(out0), _C.ptr(out))
        ctx.shard, ctx.c, ctx.n, ctx.lo, ctx.has_bias = shard, c, n, lo, bias is not None
        ctx.save_for_backward(out0, out1, beta)
        return out

    @staticmethod
    def backward(ctx, g):
        from . import _C, functional as SF
        out0, out1, beta = ctx.saved_tensors
        shard, c, n = ctx.shard, ctx.c, ctx.n
        nl, cp = out1.shape
        g = g.contiguous()
        dbeta = torch.zeros(1, dtype=torch.float32, device=g.device)
        if nl:
            part = torch.empty(_C.PARTIALS, dtype=torch.float32, device=g.device)
            _C.call("sng_pp_beta_grad", g, _C.ptr(out0), _C.ptr(out1), _C.ptr(g), g.numel(), _C.ptr(dbeta), _C.ptr(part))
        g0 = g * beta
        dw = _complete_dw(g0, shard, n, nl, c, cp)
        dbias = g.sum(0)[:c] if ctx.has_bias else None
        return g - g0, dw, g0.sum(0)[:c], dbeta, dbias, None, None, None


def _gather_row_blocks_async(local, n):
    """Start the all-gather of row blocks without waiting for it: returns (work | None, result [n, ld], buffers the caller keeps
    alive until it has waited).  The collective runs on the backend's own stream; `work.wait()` makes the current stream
    wait for it.  (NCCL only; other backends gather at once.)"""
    ws, _ = world()
    local = local.contiguous()
    if ws == 1:
        return None, local, None
    if dist.get_backend() != "nccl":
        return None, all_gather_rows(local, n), None
    r = rows_per_rank(n, ws)
    ld = local.size(1)
    pad = local if local.size(0) == r else torch.zeros(r, ld, dtype=local.dtype, device=local.device)
    if pad is not local:
        pad[: local.size(0)].copy_(local)
    out = torch.empty(ws * r, ld, dtype=local.dtype, device=local.device)
    work = dist.all_gather_into_tensor(out, pad, async_op=True)
    return work, out[:n], (pad, out)


def _complete_dw(g0, shard, n, nl, c, cp, pending=None):
    """dL/dw.weight [C, N] (complete, identical on every rank) from this shard's g0 = beta * dL/dout rows: row t of dL/dW^T
    gathers g0 over the (shifted) sources of t's in-edges -- any row of g0, hence the all-gather.  `pending` = an all-gather of
    g0 the caller already started (`_gather_row_blocks_async`), so that it overlaps the caller's kernels."""
    from . import functional as SF
    if pending is None:
        g0_all = _gather_row_blocks(g0, n)
    else:
        work, g0_all, _keep = pending
        if work is not None:
            work.wait()
    if nl:
        dwt_local = SF.spmm(g0_all.contiguous(), shard.rowptr_in, shard.col_in_shift, nl)
    else:
        dwt_local = g0.new_zeros(0, cp)
    dwt = _gather_row_blocks(dwt_local, n)                                        # [n, Cp]
    return (dwt if cp == c else dwt[:, :c]).t()


class _ShardedFusedAgg(torch.autograd.Function):
    """SNGNN++ layer on a row shard of a SYMMETRIC graph: aggregation, structural term and beta blend in ONE pass over the
    shard's in-edges (sng_edge_fwd with the fused epilogue; R: models/models.py:124-136).  Backward: scatter form of the
    aggregation backward (contribution to dL/dh of every node, reduce-scattered by AllGatherRows), dL/dW^T as above."""

    @staticmethod
    def forward(ctx, h_all, w_weight, w_bias, beta, bias, shard, n, lo, top_k, thr):
        import torch.nn.functional as F
        from . import functional as SF
        h_all = SF._check_h(h_all)
        cp = h_all.size(1)
        c = w_weight.size(0)
        fuse = (SF._padded_wt(w_weight, cp), F.pad(w_bias.detach(), (0, cp - c)).contiguous(), beta.detach().contiguous(),
                None if bias is None else F.pad(bias.detach(), (0, cp - c)).contiguous())
        train = any(ctx.needs_input_grad)
        out, sel_src, sel_w, _, sel_cnt, inv_norm, diff = SF._edge_fwd(h_all, shard, lo, int(top_k), thr, train, fuse)
        ctx.shard, ctx.c, ctx.n, ctx.lo, ctx.k, ctx.has_bias = shard, c, n, lo, int(top_k), bias is not None
        ctx.save_for_backward(h_all, sel_src, sel_w, sel_cnt, inv_norm, diff, beta)
        return out

    @staticmethod
    def backward(ctx, g):
        from . import _C
        h_all, sel_src, sel_w, sel_cnt, inv_norm, diff, beta = ctx.saved_tensors
        shard, c, n, k = ctx.shard, ctx.c, ctx.n, ctx.k
        n_total, cp = h_all.shape
        nl = shard.n
        g = g.contiguous()
        g0 = g * beta
        pending = _gather_row_blocks_async(g0, n)                 # needed only for dL/dW^T below: overlaps the aggregation backward
        g1 = (g - g0).contiguous()
        dval, dnrm, dh = torch.zeros_like(h_all), torch.zeros_like(h_all), torch.empty_like(h_all)
        _C.call("sng_edge_agg_bwd", h_all, _C.ptr(h_all), _C.ptr(inv_norm), _C.ptr(g1), n_total, nl, ctx.lo, cp, cp,
                _C.ptr(shard.rowptr_in), _C.ptr(shard.col_in), k, _C.ptr(sel_src), _C.ptr(sel_w), _C.ptr(sel_cnt),
                _C.ptr(shard.inv_deg), _C.ptr(dval), _C.ptr(dnrm), _C.ptr(dh))
        dbeta = (diff * g).sum().reshape(1)
        dw = _complete_dw(g0, shard, n, nl, c, cp, pending)
        dbias = g.sum(0)[:c] if ctx.has_bias else None
        return dh, dw, g0.sum(0)[:c], dbeta, dbias, None, None, None, None, None


def allreduce_grads(params):
    """Sum the (partial, per-shard) parameter gradients over ranks: every parameter is replicated, every rank holds the
    gradient contribution of its own rows.  Structural weights whose gradient the sharded backward already assembled on
    every rank (row blocks all-gathered, see _complete_dw) are skipped."""
    ws, _ = world()
    if ws == 1:
        return
    for p in params:
        if p.grad is not None and not getattr(p, "_sng_grad_complete", False):
            dist.all_reduce(p.grad, op=dist.ReduceOp.SUM)


def sharded_forward(model, x_local, edge_index, n, agg=None, fuse=None):
    """Row-sharded forward of an SNGNN / SNGNN_Plus / SNGNN_Plus_Plus model (replicated parameters): this rank holds the
    feature rows [lo, hi) and returns the log-probabilities of those rows.  Mirrors _SNStack.forward / the conv forwards of
    models.py (R: models/models.py:76-86,116-137,233-242,322-329) with the aggregation and the ++ fusion replaced by their
    sharded forms.  BatchNorm needs statistics over all nodes and is not supported here (the reference default is bn=False);
    neither are the non-reference options candidates='all_pairs' / denominator='selected'."""
    import torch.nn.functional as F
    from . import graph as G, models as M
    if getattr(model, "bn", False):
        raise NotImplementedError("sharded_forward: bn=True needs cross-rank batch statistics")
    ws, rank = world()
    lo, hi = shard_bounds(n, ws, rank)
    x = x_local
    convs = list(model.lins)
    for i, conv in enumerate(convs):
        base = type(conv) is M.SNConv
        plus_plus = isinstance(conv, M.SNConv_plus_plus)
        if not base and (conv.candidates != "edges" or conv.denominator != "candidates"):
            raise NotImplementedError("sharded_forward supports candidates='edges', denominator='candidates' only")
        if not base and conv.top_k <= 0:
            raise NotImplementedError("sharded_forward: top_k <= 0")
        g = G.prepare(edge_index, n, remove_self_loops=False if base else bool(conv.is_remove_self_loops), structural=plus_plus)
        h = conv._hidden_norm(x)[0]                          # lin + bias in the library's kernel (its backward is the one-pass dW / db)
        injected = agg is not None or fuse is not None
        if plus_plus:
            conv.w.weight._sng_grad_complete = not injected      # the CUDA backward assembles the complete gradient itself
        if plus_plus and not injected and g.symmetric:
            h_all = AllGatherRows.apply(h, n)
            out = _ShardedFusedAgg.apply(h_all, conv.w.weight, conv.w.bias, conv.beta, conv.bias, g.row_slice(lo, hi), n, lo,
                                         conv.top_k, conv.thr)
            out = out[:, :conv.lin.out_features]
        else:
            out, shard = edge_agg_sharded(h, g, None if base else conv.top_k, None if base else conv.thr, agg=agg)
            if plus_plus:
                out = pp_fuse_sharded(out, conv.w.weight, conv.w.bias, conv.beta, conv.bias, shard, n, lo, fuse=fuse)
                out = out[:, :conv.lin.out_features]
            else:
                out = out[:, :conv.lin.out_features]
                if conv.bias is not None:
                    out = out + conv.bias
        if i < len(convs) - 1:
            out = model.dropout(F.relu(out))
        x = out
    return F.log_softmax(x, dim=1)
