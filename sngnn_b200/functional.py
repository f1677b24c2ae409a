"""Autograd functions over the C-ABI kernels (include/sng.h).  torch is used for memory, streams and the
dense `lin` GEMM only; every sparse / selection / aggregation step is a libsng.so call.

Channel padding: the edge kernels want 16-byte feature rows, so the layer output width C is padded to
Cp = 4*ceil(C/4) by zero-padding the `lin` weights (zero columns change neither norms nor dot products).
"""
import torch
import torch.nn.functional as F

from . import _C


# Test / benchmark hook: set to a list and every EdgeAgg forward in training mode appends its (sel_src, sel_cnt) -- the
# selection lists of the layers in call order (the parity gates compare them with the oracle's rule).  None = off.
record_selection = None


def padded_channels(c):
    return (c + 3) // 4 * 4


class LinNorm(torch.autograd.Function):
    """h [N, Cp] = x W^T + b (channels zero-padded to Cp) and inv_norm = 1 / max(||h_i||, 1e-12) in ONE pass over x
    (sng_lin_norm_fwd; R: models/models.py:121-122).  Backward is the dense layer's: three library GEMMs / reductions."""

    @staticmethod
    def forward(ctx, x, weight, bias, cp):
        _C.require_cuda(x, weight, bias)
        x = x.contiguous().float()
        w = weight.contiguous()
        n, f = x.shape
        c = w.size(0)
        h = torch.empty(n, cp, dtype=torch.float32, device=x.device)
        inv = torch.empty(n, dtype=torch.float32, device=x.device)
        _C.call("sng_lin_norm_fwd", x, _C.ptr(x), n, f, f, _C.ptr(w), c, f, _C.ptr(None if bias is None else bias.contiguous()), _C.ptr(h), _C.ptr(inv))
        ctx.save_for_backward(x, w)
        ctx.has_bias, ctx.c = bias is not None, c
        ctx.mark_non_differentiable(inv)
        return h, inv

    @staticmethod
    def backward(ctx, gh, _ginv):
        x, w = ctx.saved_tensors
        dx = gh[:, :ctx.c] @ w if ctx.needs_input_grad[0] else None
        want_b = ctx.has_bias and ctx.needs_input_grad[2]
        if ctx.needs_input_grad[1] and _C.lib().sng_lin_bwd_supported(x.size(1), ctx.c):
            dw, db = lin_bwd(gh, x, ctx.c, want_b)
        else:
            g = gh[:, :ctx.c]
            dw = g.t() @ x if ctx.needs_input_grad[1] else None
            db = g.sum(0) if want_b else None
        return dx, dw, db, None


def lin_bwd(gh, x, c, want_bias=True):
    """(dW [c, f], db [c] | None) = (gh[:, :c]^T x, column sums of gh[:, :c]) in one streaming pass (sng_lin_bwd): the weight
    gradient of `self.lin` (R: models/models.py:121).  The library GEMM runs this N-long reduction with a [c, f] output as one
    skinny kernel at a fraction of the memory bandwidth; fixed-order partial sums make the result bit-reproducible."""
    _C.require_cuda(gh, x)
    if gh.stride(1) != 1:
        gh = gh.contiguous()
    if x.stride(1) != 1:
        x = x.contiguous()
    n, f = x.shape
    dw = torch.empty(c, f, dtype=torch.float32, device=x.device)
    db = torch.empty(c, dtype=torch.float32, device=x.device) if want_bias else None
    nb = _C.lib().sng_lin_bwd_workspace_bytes(n, f, c)
    ws = torch.empty(nb, dtype=torch.uint8, device=x.device)
    _C.call("sng_lin_bwd", x, _C.ptr(gh), gh.stride(0), _C.ptr(x), x.stride(0), n, f, c, _C.ptr(dw), f, _C.ptr(db), _C.ptr(ws), nb)
    return dw, db


def lin_norm(x, weight, bias, cp):
    """(h [N, Cp], inv_norm [N] | None): the fused kernel when the shape is supported, else library GEMM (inv_norm = None: the
    aggregation call then computes it)."""
    if x.is_cuda and _C.lib().sng_lin_norm_supported(x.size(1), weight.size(0)):
        return LinNorm.apply(x, weight, bias, cp)
    return linear_padded(x, weight, bias, cp), None


def linear_padded(x, weight, bias, cp):
    """h = x @ W^T + b with the output width zero-padded to `cp` (R: models/models.py:121,237,324)."""
    c = weight.size(0)
    if cp != c:
        weight = F.pad(weight, (0, 0, 0, cp - c))
        bias = None if bias is None else F.pad(bias, (0, cp - c))
    return F.linear(x, weight, bias)


def _check_h(h):
    _C.require_cuda(h)
    if h.dtype != torch.float32 or h.dim() != 2 or h.size(1) % 4 != 0:
        raise RuntimeError("expected a float32 [N, 4m] tensor")
    return h.contiguous()


def _edge_fwd(h_all, graph, row_offset, k, thr, train, fuse=None, want_q=False, inv_norm=None):
    """sng_edge_fwd on the target rows of `graph` (a PreparedGraph or a row shard of one).  fuse = (wt [N, Cp], b_w [Cp],
    beta [1], bias [Cp] | None) gathers the structural term and blends in the same pass (symmetric graphs only)."""
    c = h_all.size(1)
    n = graph.n
    dev = h_all.device
    out = torch.empty(n, c, dtype=torch.float32, device=dev)
    sel_src = sel_w = sel_cnt = sel_q = diff = None
    if train and k > 0:
        sel_src = torch.empty(n, k, dtype=torch.int32, device=dev)
        sel_w = torch.empty(n, k, dtype=torch.float32, device=dev)
        sel_cnt = torch.empty(n, dtype=torch.int32, device=dev)
        if want_q:
            sel_q = torch.empty(n, k, dtype=torch.int32, device=dev)
    inv_ready = inv_norm is not None                      # produced with h by sng_lin_norm_fwd
    if not inv_ready:
        inv_norm = torch.empty(h_all.size(0), dtype=torch.float32, device=dev)
    wt = bw = beta = bias = None
    if fuse is not None:
        wt, bw, beta, bias = fuse
        if train:
            diff = torch.empty_like(out)
    tab, ws, wbytes = graph.chunk_tab, None, 0
    n_chunks = -1 if tab is None else int(tab.size(0))
    if n_chunks > 0 and c <= 32:
        wbytes = _C.lib().sng_edge_fwd_workspace_bytes(n_chunks, c, k)
        ws = torch.empty(wbytes, dtype=torch.uint8, device=dev)
    n_lrows = 0 if tab is None else int(graph.lrows.numel())
    n_hub = 0 if graph.rows_hub is None else int(graph.rows_hub.numel())
    _C.call("sng_edge_fwd", h_all, _C.ptr(h_all), h_all.size(0), n, row_offset, c, c, _C.ptr(graph.rowptr_in), _C.ptr(graph.col_in),
            _C.ptr(graph.tpos if want_q else None), _C.ptr(tab), n_chunks, _C.ptr(graph.lrows), _C.ptr(graph.lrow_ptr), n_lrows,
            _C.ptr(graph.rows_hub), n_hub, _C.ptr(ws), wbytes, k,
            float(thr if thr is not None else 0.0), _C.ptr(out), c, _C.ptr(sel_src), _C.ptr(sel_w), _C.ptr(sel_q), _C.ptr(sel_cnt),
            _C.ptr(inv_norm), int(inv_ready), _C.ptr(wt), c, _C.ptr(bw), _C.ptr(beta), _C.ptr(bias), _C.ptr(diff))
    return out, sel_src, sel_w, sel_q, sel_cnt, inv_norm, diff


def _padded_wt(w_weight, cp):
    """W^T [N, Cp] of the structural weight [C, N].  The modules keep that parameter in transposed storage
    (models.SNConv_plus_plus), so this is a view of the parameter itself when C is a multiple of 4, and a small padded copy
    otherwise; a parameter in plain [C, N] storage costs a transpose copy per call.  Nothing is cached."""
    c = w_weight.size(0)
    wt = w_weight.detach().t()
    if cp != c:
        wt = F.pad(wt, (0, cp - c))
    return wt.contiguous()


class EdgeAgg(torch.autograd.Function):
    """The similarity-navigated aggregation of one layer: out_1 of R: models/models.py:132 / :239 / :326 and -- when the
    structural parameters are given and the graph's in-lists equal its out-lists -- the whole of
    R: models/models.py:124-136 (out = beta (A W^T + b_w) + (1 - beta) out_1 + bias) in the same pass over the edges.
    Forward sng_edge_fwd, backward sng_edge_bwd (two gather passes, no float atomics, bit-reproducible)."""

    @staticmethod
    def forward(ctx, h, graph, top_k, thr, w_weight, w_bias, beta, bias, inv_norm=None):
        h = _check_h(h)
        n, cp = h.shape
        if n != graph.n:
            raise RuntimeError(f"h has {n} rows but the graph has {graph.n} nodes")
        k = int(top_k) if top_k is not None else 0
        fused = w_weight is not None
        train = any(ctx.needs_input_grad)
        fuse = None
        if fused:
            _C.require_cuda(h, w_weight, w_bias, beta, bias)
            c = w_weight.size(0)
            if w_weight.size(1) != n:
                raise RuntimeError(f"w.weight is [{c},{w_weight.size(1)}] but the graph has {n} nodes "
                                   "(R builds w = Linear(num_nodes, out_channels), models.py:95)")
            if not graph.symmetric:
                raise RuntimeError("the fused SNGNN++ pass needs a graph whose in-lists equal its out-lists")
            fuse = (_padded_wt(w_weight, cp), F.pad(w_bias.detach(), (0, cp - c)).contiguous(), beta.detach().contiguous(),
                    None if bias is None else F.pad(bias.detach(), (0, cp - c)).contiguous())
            ctx.c = c
        out, sel_src, sel_w, sel_q, sel_cnt, inv_norm, diff = _edge_fwd(h, graph, 0, k, thr, train, fuse, want_q=True, inv_norm=inv_norm)
        ctx.graph, ctx.k, ctx.fused, ctx.has_bias = graph, k, fused, bias is not None
        if record_selection is not None and sel_cnt is not None:
            record_selection.append((sel_src, sel_cnt))
        ctx.save_for_backward(h, sel_src, sel_w, sel_q, sel_cnt, inv_norm, diff, beta if fused else None)
        if sel_cnt is not None:
            ctx.mark_non_differentiable(sel_src, sel_w, sel_cnt)
        return out, sel_src, sel_w, sel_cnt

    @staticmethod
    def backward(ctx, g, *_):
        h, sel_src, sel_w, sel_q, sel_cnt, inv_norm, diff, beta = ctx.saved_tensors
        graph, k, fused = ctx.graph, ctx.k, ctx.fused
        cp = h.size(1)
        g = g.contiguous()
        dh, dwt, dbeta = edge_bwd(h, inv_norm, g, graph, k, sel_src, sel_w, sel_q, sel_cnt, beta if fused else None, diff)
        if not fused:
            return dh, None, None, None, None, None, None, None, None
        c = ctx.c
        gsum = g.sum(0)[:c]
        dw = (dwt if cp == c else dwt[:, :c]).t()           # [C, N] in the parameter's own (transposed) layout
        return dh, None, None, None, dw, gsum * beta, dbeta, (gsum if ctx.has_bias else None), None


def edge_bwd(h, inv_norm, g, graph, k, sel_src, sel_w, sel_q, sel_cnt, beta=None, diff=None):
    """sng_edge_bwd: dL/dh [N, Cp] of the aggregation -- two gather passes, no float atomics -- and, with `beta` (the fused
    SNGNN++ epilogue), dL/dW^T [N, Cp] and dL/dbeta.  `g` = dL/dout."""
    n, cp = h.shape
    dev = h.device
    fused = beta is not None
    coef = torch.empty(max(graph.num_edges, 1) * 2, dtype=torch.float32, device=dev)
    dnt = torch.empty_like(h)
    dh = torch.empty_like(h)
    dbeta = dwt = part = None
    if fused:
        dbeta = torch.empty(1, dtype=torch.float32, device=dev)
        part = torch.empty(_C.PARTIALS, dtype=torch.float32, device=dev)
        dwt = torch.empty(n, cp, dtype=torch.float32, device=dev)
    tab = graph.chunk_tab_out if cp <= 32 else None
    n_chunks = -1 if tab is None else int(tab.size(0))
    cpart = None
    if n_chunks > 0:
        cg = 4
        while cg < cp:
            cg *= 2
        cpart = torch.empty(n_chunks * 3 * cg, dtype=torch.float32, device=dev)
    _C.call("sng_edge_bwd", h, _C.ptr(h), _C.ptr(inv_norm), _C.ptr(g), n, cp, cp, cp, _C.ptr(graph.rowptr_in), _C.ptr(graph.col_in),
            _C.ptr(graph.tpos), _C.ptr(graph.rowptr_out), _C.ptr(graph.col_out), graph.src_shift, graph.num_edges, k,
            _C.ptr(sel_src), _C.ptr(sel_w), _C.ptr(sel_q), _C.ptr(sel_cnt), _C.ptr(beta), _C.ptr(diff if fused else None), cp,
            _C.ptr(dbeta), _C.ptr(coef), _C.ptr(dnt), _C.ptr(part), _C.ptr(dh), _C.ptr(dwt), cp,
            _C.ptr(tab), n_chunks, _C.ptr(graph.lrows_out if tab is not None else None),
            _C.ptr(graph.lrow_ptr_out if tab is not None else None), 0 if tab is None else int(graph.lrows_out.numel()), _C.ptr(cpart))
    return dh, dwt, dbeta


def edge_topk_agg_rows(h_all, shard, row_offset, top_k=None, thr=None):
    """Forward-only K2 on a row shard: targets [row_offset, row_offset + shard.n) of `h_all` (all nodes, e.g. after an
    all-gather), `shard` = PreparedGraph.row_slice(lo, hi).  Returns (out [shard.n, C], sel_src, sel_w, sel_cnt)."""
    h_all = _check_h(h_all)
    out, sel_src, sel_w, _, sel_cnt, _, _ = _edge_fwd(h_all, shard, int(row_offset), int(top_k) if top_k is not None else 0, thr, True)
    return out, sel_src, sel_w, sel_cnt


class ShardedEdgeTopkAgg(torch.autograd.Function):
    """K2 / K2b on a ROW SHARD: forward computes out_1 for target rows [row_offset, row_offset + shard.n) from `h_all` (all
    nodes); backward returns this shard's contribution to dL/dh_all [N, C] -- contributions of different shards add
    (the caller reduce-scatters them, see dist.AllGatherRows).  A shard has no transpose index, so its
    backward is the scatter form (sng_edge_agg_bwd)."""

    @staticmethod
    def forward(ctx, h_all, shard, row_offset, top_k, thr):
        h_all = _check_h(h_all)
        k = int(top_k) if top_k is not None else 0
        out, sel_src, sel_w, _, sel_cnt, inv_norm, _ = _edge_fwd(h_all, shard, int(row_offset), k, thr, ctx.needs_input_grad[0])
        ctx.shard, ctx.k, ctx.row_offset = shard, k, int(row_offset)
        ctx.save_for_backward(h_all, sel_src, sel_w, sel_cnt, inv_norm)
        return out

    @staticmethod
    def backward(ctx, g):
        h_all, sel_src, sel_w, sel_cnt, inv_norm = ctx.saved_tensors
        shard, k = ctx.shard, ctx.k
        n_total, c = h_all.shape
        g = g.contiguous()
        dval, dnrm, dh = torch.zeros_like(h_all), torch.zeros_like(h_all), torch.empty_like(h_all)
        _C.call("sng_edge_agg_bwd", h_all, _C.ptr(h_all), _C.ptr(inv_norm), _C.ptr(g), n_total, shard.n, ctx.row_offset, c, c,
                _C.ptr(shard.rowptr_in), _C.ptr(shard.col_in), k, _C.ptr(sel_src), _C.ptr(sel_w), _C.ptr(sel_cnt),
                _C.ptr(shard.inv_deg), _C.ptr(dval), _C.ptr(dnrm), _C.ptr(dh))
        return dh, None, None, None, None


def edge_topk_agg(h, graph, top_k=None, thr=None, return_selection=False, structural=None, inv_norm=None):
    """out_1 (structural=None) or the fused SNGNN++ layer output (structural = (w.weight, w.bias, beta, bias)).
    inv_norm = 1 / max(||h_i||, 1e-12) when the caller already has it (lin_norm), else it is computed here."""
    w_weight, w_bias, beta, bias = structural if structural is not None else (None, None, None, None)
    out, sel_src, sel_w, sel_cnt = EdgeAgg.apply(h, graph, top_k, thr, w_weight, w_bias, beta, bias, inv_norm)
    if return_selection:
        return out, (sel_src, sel_w, sel_cnt)
    return out


class ListAgg(torch.autograd.Function):
    """Mean aggregation over an explicit neighbour list (all-pairs mode): forward sng_list_agg_fwd, backward K2b."""

    @staticmethod
    def forward(ctx, h, idx, sim, cnt, inv_denom):
        h = _check_h(h)
        n, c = h.shape
        out = torch.empty(idx.size(0), c, dtype=h.dtype, device=h.device)
        _C.call("sng_list_agg_fwd", h, _C.ptr(h), idx.size(0), c, c, idx.size(1), _C.ptr(idx), _C.ptr(sim), _C.ptr(cnt),
                _C.ptr(inv_denom), _C.ptr(out), c)
        ctx.save_for_backward(h, idx, sim, cnt, inv_denom)
        return out

    @staticmethod
    def backward(ctx, g):
        h, idx, sim, cnt, inv_denom = ctx.saved_tensors
        n, c = h.shape
        if idx.size(0) != n:
            raise RuntimeError("backward through a row-sharded list aggregation is not supported")
        g = g.contiguous()
        dval, dnrm, dh = torch.zeros_like(h), torch.zeros_like(h), torch.empty_like(h)
        _, _, inv_norm = rownorm(h, want_f32=False, want_inv=True)
        _C.call("sng_edge_agg_bwd", h, _C.ptr(h), _C.ptr(inv_norm), _C.ptr(g), n, n, 0, c, c, None, None, idx.size(1), _C.ptr(idx), _C.ptr(sim),
                _C.ptr(cnt), _C.ptr(inv_denom), _C.ptr(dval), _C.ptr(dnrm), _C.ptr(dh))
        return dh, None, None, None, None


def spmm(x, rowptr, col, n_rows, val=None, rowscale=None, bias=None):
    """K3: out[i] = rowscale[i] * sum_e val[e] x[col[e]] + bias (no autograd; used inside backward passes)."""
    x = _check_h(x)
    c = x.size(1)
    out = torch.empty(n_rows, c, dtype=x.dtype, device=x.device)
    _C.call("sng_spmm_fwd", x, _C.ptr(x), n_rows, c, c, _C.ptr(rowptr), _C.ptr(col), _C.ptr(val), _C.ptr(rowscale),
            _C.ptr(bias), _C.ptr(out), c)
    return out


class PPFuse(torch.autograd.Function):
    """out = beta*(A @ W^T + b_w) + (1-beta)*out_1 (+bias)  --  R: models/models.py:124-136 (K4)."""

    @staticmethod
    def forward(ctx, out1, w_weight, w_bias, beta, bias, graph):
        out1 = _check_h(out1)
        n, cp = out1.shape
        c = w_weight.size(0)
        if w_weight.size(1) != n:
            raise RuntimeError(f"w.weight is [{c},{w_weight.size(1)}] but the graph has {n} nodes "
                               "(R builds w = Linear(num_nodes, out_channels), models.py:95)")
        wt = _padded_wt(w_weight, cp)                                             # [N, Cp]
        bw = F.pad(w_bias.detach(), (0, cp - c)).contiguous()
        bb = None if bias is None else F.pad(bias.detach(), (0, cp - c)).contiguous()
        out0 = torch.empty_like(out1)
        out = torch.empty_like(out1)
        _C.call("sng_pp_fuse_fwd", out1, _C.ptr(wt), n, cp, cp, _C.ptr(graph.rowptr_out), _C.ptr(graph.col_out), _C.ptr(bw),
                _C.ptr(beta), _C.ptr(out1), _C.ptr(bb), _C.ptr(out0), _C.ptr(out))
        ctx.graph, ctx.c, ctx.has_bias = graph, c, bias is not None
        ctx.save_for_backward(out0, out1, beta)
        return out

    @staticmethod
    def backward(ctx, g):
        out0, out1, beta = ctx.saved_tensors
        graph, c = ctx.graph, ctx.c
        n, cp = out1.shape
        g = g.contiguous()
        dbeta = torch.empty(1, dtype=torch.float32, device=g.device)
        part = torch.empty(_C.PARTIALS, dtype=torch.float32, device=g.device)
        g0, dout1, dwt = torch.empty_like(g), torch.empty_like(g), torch.empty_like(g)
        # dL/dW^T = A^T g0: row t gathers g0 over the (shifted) sources of t's in-edges
        _C.call("sng_pp_fuse_bwd", g, _C.ptr(out0), _C.ptr(out1), _C.ptr(g), _C.ptr(beta), n, cp, cp, _C.ptr(graph.rowptr_in),
                _C.ptr(graph.col_in_shift), _C.ptr(dbeta), _C.ptr(part), _C.ptr(g0), _C.ptr(dout1), _C.ptr(dwt))
        dw = (dwt if cp == c else dwt[:, :c]).t()
        db_w = g0.sum(0)[:c]
        dbias = g.sum(0)[:c] if ctx.has_bias else None
        return dout1, dw, db_w, dbeta, dbias, None


def rownorm(x, want_f32=True, want_f16=False, f16_ld=None, want_inv=False):
    """K0: F.normalize(x, p=2, dim=-1) (R: models/models.py:122).  Returns (xhat_f32, xhat_f16, inv_norm)."""
    _C.require_cuda(x)
    x = x.contiguous().float()
    n, d = x.shape
    xf = torch.empty_like(x) if want_f32 else None
    ldh = f16_ld or (d + 15) // 16 * 16
    xh = torch.empty(n, ldh, dtype=torch.float16, device=x.device) if want_f16 else None
    inv = torch.empty(n, dtype=torch.float32, device=x.device) if want_inv else None
    _C.call("sng_rownorm_f32", x, _C.ptr(x), n, d, d, _C.ptr(xf), d, _C.ptr(xh), ldh, _C.ptr(inv))
    return xf, xh, inv


def sddmm_dot(xhat, a, b):
    """s[e] = <xhat[a[e]], xhat[b[e]]> (R: SimGFAToolbox/dense.py:160-162)."""
    _C.require_cuda(xhat, a, b)
    xhat = xhat.contiguous()
    a = a.to(torch.int32).contiguous()
    b = b.to(torch.int32).contiguous()
    s = torch.empty(a.numel(), dtype=torch.float32, device=xhat.device)
    _C.call("sng_sddmm_dot", xhat, _C.ptr(xhat), xhat.size(0), xhat.size(1), xhat.size(1), _C.ptr(a), _C.ptr(b), a.numel(), _C.ptr(s))
    return s


class _NllLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logp, y):
        _C.require_cuda(logp, y)
        logp = logp.contiguous()
        y = y.contiguous().long()
        n, c = logp.shape
        out = torch.empty(2, dtype=torch.float32, device=logp.device)             # loss, count
        part = torch.empty(_C.PARTIALS, dtype=torch.float32, device=logp.device)
        _C.call("sng_nll_loss_fwd", logp, _C.ptr(logp), n, c, c, _C.ptr(y), _C.ptr(out), _C.ptr(out[1:]), _C.ptr(part))
        ctx.save_for_backward(y, out)
        ctx.shape = (n, c)
        return out[0]

    @staticmethod
    def backward(ctx, g):
        y, out = ctx.saved_tensors
        n, c = ctx.shape
        d = torch.empty(n, c, dtype=torch.float32, device=y.device)
        g = g.contiguous().float().reshape(1)
        _C.call("sng_nll_loss_bwd", y, n, c, c, _C.ptr(y), _C.ptr(g), _C.ptr(out[1:]), _C.ptr(d))
        return d, None


def nll_loss(logp, y, mask=None):
    """Mean negative log-likelihood of `logp` [N, C] (the models' log_softmax output) against labels `y` [N] -- what
    R: train.py:81 computes as F.nll_loss(out[mask], y[mask]).  `mask` (bool [N]) keeps the masked form in ONE pass: masked-out
    rows get the label -1 instead of being gathered away.  Fixed-order reductions: bit-reproducible, forward and backward."""
    if mask is not None:
        y = torch.where(mask, y, torch.full_like(y, -1))
    return _NllLoss.apply(logp, y)
