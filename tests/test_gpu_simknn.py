"""GPU parity of the all-pairs similarity-kNN builder (K1: tcgen05 stage 1, FP32 rescore, exact fallback)."""
import pytest
import torch

from parity import compare_lists, check_tie_order

pytestmark = pytest.mark.gpu


@pytest.fixture
def debug_env():
    """The planner's SNG_KNN_* overrides are read only while a test switches them on (include/sng.h: sng_set_debug_env)."""
    from sngnn_b200 import _C
    _C.lib().sng_set_debug_env(1)
    yield
    _C.lib().sng_set_debug_env(0)

DEV = "cuda"


def _features(n, d, kind, seed):
    from sngnn_b200 import synth
    return synth.make_features(n, d, kind, seed=seed, dup_frac=0.01, zero_rows=3)


def _oracle(x, k, thr, remove_self, q_lo=0, q_hi=None):
    from oracle import sn_ref
    return sn_ref.simknn_allpairs(x.double(), k, thr, remove_self, q_lo, q_hi, dtype=torch.float64)


def _score64(x):
    from oracle import sn_ref
    n64 = sn_ref.rownorm(x.double())
    return lambda r, j: (n64[r] * n64[j]).sum(-1)


@pytest.mark.parametrize("n,d,ew,ns", [(1000, 64, 1, 1), (1000, 64, 2, 1), (3000, 65, 4, 3), (2500, 128, 4, 2), (1500, 269, 0, 0),
                                       (1300, 512, 0, 0), (700, 16, 2, 1), (900, 40, 0, 0), (5000, 65, 0, 0), (257, 65, 4, 1), (900, 24, 0, 0), (800, 90, 4, 1), (600, 200, 0, 0),
                                       # d > ~640: the query block is streamed through the ring with B (Stage1Params::stream_a)
                                       (1500, 768, 0, 0), (1200, 1024, 0, 0), (900, 2325, 0, 0)])
def test_stage1_candidates_contain_true_topk(n, d, ew, ns):
    """Tensor-core stage.  A candidate is a column TRIPLE {c, c+1, c+2} (two columns when c % 32 == 30) scored with the
    maximum FP16 score of its columns.  Every true top-10 neighbour must lie inside a kept triple or under the reported
    drop bound, and the kept scores must be within the proven error bound of the exact cosines."""
    from sngnn_b200 import simknn
    x = _features(n, d, "normal", seed=n + d)
    ci, cv, cm, xf, xh = simknn.stage1_candidates(x.to(DEV), 16, thr_lo=-2.0, remove_self=True, force_ew=ew, force_nsplit=ns)
    torch.cuda.synchronize()
    ci, cv, cm = ci.cpu().long().reshape(n, -1), cv.cpu().reshape(n, -1), cm.cpu()
    score = _score64(x)
    rows = torch.arange(n)[:, None].expand_as(ci)
    best = torch.full(ci.shape, float("-inf"), dtype=torch.float64)
    for e in range(3):
        col = ci + e
        ok = (ci >= 0) & (col < n) & (col != rows) & ((e < 2) | (ci % 32 != 30))
        v = torch.full(ci.shape, float("-inf"), dtype=torch.float64)
        v[ok] = score(rows[ok], col[ok])
        best = torch.maximum(best, v)
    kept = ci >= 0
    assert torch.isfinite(best[kept]).all()
    assert (cv[kept].double() - best[kept]).abs().max() < 1.1e-3
    idx_ref, _, cnt_ref = _oracle(x, 10, -2.0, True)
    zero_rows = (x.abs().sum(1) == 0)
    for r in range(n):
        if zero_rows[r]:
            continue                                     # all scores tie at 0: any candidate set is legal for stage 1
        want = set(idx_ref[r, :cnt_ref[r]].tolist())
        have = set()
        for c in ci[r][ci[r] >= 0].tolist():
            have.update(range(c, c + (2 if c % 32 == 30 else 3)))
        missing = want - have
        if missing:                                      # anything dropped must be covered by the reported drop bound
            vals = score(torch.full((len(missing),), r), torch.tensor(sorted(missing)))
            assert (vals <= cm[r].max().item() + 1.1e-3).all(), (r, missing)


@pytest.mark.parametrize("n,d,k,thr,rs,kind", [(2000, 65, 10, -1.0, True, "normal"), (2000, 65, 10, 0.9, True, "clustered"),
                                                (3000, 128, 10, 0.0, False, "clustered"), (1200, 48, 50, -1.0, True, "clustered"),
                                                (2277, 2325 // 8, 10, 0.3, True, "binary"), (1000, 269, 5, 0.5, True, "clustered"),
                                                (600, 8, 64, -1.0, False, "normal"), (130, 33, 10, -1.0, True, "normal"),
                                                # wide features (streamed query block): the Chameleon shape with its raw 2,325 features
                                                (2277, 2325, 10, 0.3, True, "binary"), (1500, 800, 10, -1.0, True, "normal"),
                                                (1000, 1500, 5, 0.0, False, "clustered")])
def test_build_matches_oracle(n, d, k, thr, rs, kind):
    from sngnn_b200 import simknn
    x = _features(n, d, kind, seed=7 * n + d)
    idx, sim, cnt, nfb = simknn.build_knn(x.to(DEV), k, thr, rs, return_fallback=True)
    torch.cuda.synchronize()
    idx_ref, sim_ref, cnt_ref = _oracle(x, k, thr, rs)
    res = compare_lists(idx, cnt, idx_ref, cnt_ref, _score64(x), thr)
    print(f"n={n} d={d} k={k} thr={thr}: {res} fallback_rows={int(nfb[0])} retry_rows={int(nfb[1])}")
    assert res["out_of_band"] == 0, res
    assert check_tie_order(idx, sim, cnt)
    keep = torch.arange(k)[None, :] < cnt.cpu()[:, None]
    same = (idx.cpu().long() == idx_ref) & keep
    tol = 2e-6 if d <= 512 else 1e-5                      # FP32 summation error of a d-term dot product (the contract is 1e-5)
    assert (sim.cpu().double()[same] - sim_ref[same]).abs().max() < tol if same.any() else True


@pytest.mark.parametrize("n,d", [(1500, 40), (20000, 24)])
def test_massive_ties_go_through_exact_fallback(n, d):
    """A handful of distinct feature rows, each repeated hundreds of times: every row ties at cosine 1 with far more columns than
    any candidate list holds, so stage 2 cannot prove it and the exact stage-3 scans (parallel waves + serial remainder) must
    reproduce the index tie-break.  All-zero rows (a fifth of them) tie at 0 everywhere and are resolved by rule in stage 2."""
    from sngnn_b200 import simknn
    g = torch.Generator().manual_seed(n)
    pats = torch.randn(5, d, generator=g)
    pats[0] = 0
    x = pats[torch.randint(0, 5, (n,), generator=g)]
    k, thr = 10, -1.0
    idx, sim, cnt, nfb = simknn.build_knn(x.to(DEV), k, thr, True, return_fallback=True)
    torch.cuda.synchronize()
    idx_ref, sim_ref, cnt_ref = _oracle(x, k, thr, True)
    res = compare_lists(idx, cnt, idx_ref, cnt_ref, _score64(x), thr)
    print(f"ties n={n}: {res} fallback_rows={int(nfb[0])} retry_rows={int(nfb[1])}")
    assert int(nfb[0]) > n // 4
    assert res["out_of_band"] == 0, res
    assert check_tie_order(idx, sim, cnt)
    zero = (x.abs().sum(1) == 0)
    assert torch.equal(idx.cpu().long()[zero], idx_ref[zero]) and bool((sim.cpu()[zero] == 0).all())


@pytest.mark.parametrize("thr", [0.0, 0.5])
def test_sparse_binary_rows_with_zero_rows(thr):
    """Sparse 0/1 features (a third of the rows all-zero, most scores exactly 0): lists identical to the oracle's."""
    from sngnn_b200 import simknn
    n, d, k = 1500, 40, 10
    x = (torch.rand(n, d, generator=torch.Generator().manual_seed(n)) < 0.03).float()
    idx, sim, cnt, nfb = simknn.build_knn(x.to(DEV), k, thr, True, return_fallback=True)
    idx_ref, sim_ref, cnt_ref = _oracle(x, k, thr, True)
    res = compare_lists(idx, cnt, idx_ref, cnt_ref, _score64(x), thr)
    assert res["out_of_band"] == 0 and res["exact"] == n, res
    assert check_tie_order(idx, sim, cnt)


def test_build_row_sharded_equals_full():
    from sngnn_b200 import simknn
    n, d, k = 3000, 65, 10
    x = _features(n, d, "clustered", seed=3).to(DEV)
    full = simknn.build_knn(x, k, 0.2, True)
    parts = [simknn.build_knn(x, k, 0.2, True, q_lo=lo, q_hi=hi) for lo, hi in ((0, 1000), (1000, 1001), (1001, 3000))]
    for t in range(3):
        assert torch.equal(full[t], torch.cat([p[t] for p in parts]))


def test_allpairs_model_mode_matches_oracle():
    """SNConv_plus(candidates='all_pairs') == the reference conv run on the complete graph (its own oracle)."""
    import sngnn_b200.models as M
    from oracle import sn_ref
    torch.manual_seed(0)
    n, fd, c, k, thr = 700, 24, 8, 5, 0.1
    x = torch.randn(n, fd)
    conv = M.SNConv_plus(fd, c, n, k, thr, True, candidates="all_pairs").to(DEV)
    out = conv(x.to(DEV), torch.zeros(2, 0, dtype=torch.long, device=DEV))
    src = torch.arange(n).repeat(n); dst = torch.arange(n).repeat_interleave(n)
    keep = src != dst
    ei = torch.stack([src[keep], dst[keep]])
    h = torch.nn.functional.linear(x, conv.lin.weight.detach().cpu(), conv.lin.bias.detach().cpu())
    ref = sn_ref.sn_aggregate(h, ei, k, thr)
    torch.testing.assert_close(out.detach().cpu(), ref, rtol=1e-4, atol=1e-6)


@pytest.mark.parametrize("n,d,stride,ew", [(20000, 65, 2, 4), (17000, 128, 2, 4), (33000, 200, 4, 2), (16500, 400, 2, 1)])
def test_seed_pass_group_maxima(n, d, stride, ew):
    """SEED mode of the tensor-core kernel: seeds[r, g] = max FP16 score of row r over the sampled columns of group g
    (sample = every stride-th database row; group of sample column c = ((c>>8)&1)*8 + ((c&255)>>5))."""
    from sngnn_b200 import simknn
    x = _features(n, d, "clustered", seed=n)
    xf, xh = simknn.normalize_operands(x.to(DEV))
    nq = 700
    seeds = simknn.seed_pass(xh[:nq], xh, d, stride, ew).cpu()
    torch.cuda.synchronize()
    xs = xh[::stride].float()
    sc = (xh[:nq].float() @ xs.t()).cpu()                        # products of FP16 values, FP32 accumulation
    cs = torch.arange(xs.size(0))
    grp = ((cs >> 8) & 1) * 8 + ((cs & 255) >> 5)
    ref = torch.full((nq, 16), float("-inf"))
    for g in range(16):
        if (grp == g).any():
            ref[:, g] = sc[:, grp == g].max(1).values
    assert torch.isfinite(seeds).all() == torch.isfinite(ref).all()
    fin = torch.isfinite(ref)
    assert (seeds[fin] - ref[fin]).abs().max() < 2e-4            # summation order of the tensor cores vs torch


@pytest.mark.parametrize("n,d,k,thr,rs,kind,stride,q", [(20000, 65, 10, -1.0, True, "normal", 2, 0), (20000, 65, 10, 0.5, True, "clustered", 2, 0),
                                                         (36000, 128, 5, -1.0, False, "clustered", 4, 0), (20000, 40, 10, -1.0, True, "normal", 2, 1),
                                                         (24000, 200, 10, 0.2, True, "clustered", 2, 0), (70000, 300, 50, -1.0, True, "clustered", 32, 0)])
def test_seeded_build_matches_oracle(n, d, k, thr, rs, kind, stride, q, monkeypatch, debug_env):
    """Full build with threshold seeding switched on (SNG_KNN_SEED_S shrinks the stride so it engages at test sizes;
    q = 1 makes the seeded threshold far too aggressive, so the proof must send many rows to the exact scan)."""
    from sngnn_b200 import simknn
    monkeypatch.setenv("SNG_KNN_SEED_S", str(stride))
    if q:
        monkeypatch.setenv("SNG_KNN_SEED_Q", str(q))
    x = _features(n, d, kind, seed=n + d + k)
    plan = simknn.build_plan(n, n, d, k)
    assert plan["seed_stride"] == stride and plan["seed_q"] >= 1, plan
    idx, sim, cnt, nfb = simknn.build_knn(x.to(DEV), k, thr, rs, return_fallback=True)
    torch.cuda.synchronize()
    rows = 1500                                                  # oracle on a slab of query rows (all columns)
    sl = slice(n // 2, n // 2 + rows)
    idx_ref, sim_ref, cnt_ref = _oracle(x, k, thr, rs, sl.start, sl.stop)
    score = _score64(x)
    res = compare_lists(idx[sl], cnt[sl], idx_ref, cnt_ref, lambda r, j: score(r + sl.start, j), thr)
    print(f"seeded n={n} d={d} k={k} thr={thr} plan={plan}: {res} fallback_rows={int(nfb[0])} retry_rows={int(nfb[1])}")
    assert res["out_of_band"] == 0, res
    assert check_tie_order(idx, sim, cnt)
    if q == 1:
        assert int(nfb[1]) > 0                                      # rows that needed the retry pass


@pytest.mark.parametrize("n,d,k", [(40000, 65, 10), (20000, 128, 50), (9000, 65, 10), (30000, 512, 50), (26000, 512, 10)])
def test_default_plan_build_matches_oracle(n, d, k):
    """The plan the library picks by itself (seed stride / quantile adapted to the database size and top_k)."""
    from sngnn_b200 import simknn
    x = _features(n, d, "clustered", seed=3 * n + d)
    plan = simknn.build_plan(n, n, d, k)
    idx, sim, cnt, nfb = simknn.build_knn(x.to(DEV), k, 0.0, True, return_fallback=True)
    torch.cuda.synchronize()
    rows = 1200
    idx_ref, sim_ref, cnt_ref = _oracle(x, k, 0.0, True, 0, rows)
    res = compare_lists(idx[:rows], cnt[:rows], idx_ref, cnt_ref, _score64(x), 0.0)
    print(f"default plan n={n} d={d} k={k} plan={plan}: {res} fallback_rows={int(nfb[0])} retry_rows={int(nfb[1])}")
    assert res["out_of_band"] == 0, res
    assert check_tie_order(idx, sim, cnt)


def test_build_without_retry_pass_matches_oracle(monkeypatch, debug_env):
    """SNG_KNN_NORETRY: unproven rows go straight to the exact scan (the path every row took before the retry pass existed)."""
    from sngnn_b200 import simknn
    monkeypatch.setenv("SNG_KNN_NORETRY", "1")
    monkeypatch.setenv("SNG_KNN_CAND", "10")                     # no spare slot: many rows cannot be proven
    n, d, k = 6000, 65, 10
    x = _features(n, d, "clustered", seed=11)
    x[100:140] = x[100]                                          # 40 identical rows: 39 ties at cosine 1 against a 10-slot list
    idx, sim, cnt, nfb = simknn.build_knn(x.to(DEV), k, 0.0, True, return_fallback=True)
    torch.cuda.synchronize()
    idx_ref, sim_ref, cnt_ref = _oracle(x, k, 0.0, True)
    res = compare_lists(idx, cnt, idx_ref, cnt_ref, _score64(x), 0.0)
    print(f"no retry: {res} fallback_rows={int(nfb[0])}")
    assert res["out_of_band"] == 0, res
    assert int(nfb[0]) > 0 and int(nfb[0]) == int(nfb[1])
    assert check_tie_order(idx, sim, cnt)


def test_build_inside_a_cuda_graph_capture():
    """The build reads the flagged-row count back (one host synchronisation) to size its retry rounds; on a capturing stream it
    must not synchronise: one fixed retry round with the count on the device.  Captured + replayed lists == eager lists."""
    from sngnn_b200 import simknn
    n, d, k = 4000, 65, 10
    x = _features(n, d, "clustered", seed=5)
    x[200:500] = x[200]                                           # 300 identical rows: more ties than a main-pass list holds -> flagged
    xf, xh = simknn.normalize_operands(x.to(DEV))
    eager = simknn.build_knn_normalized(xf, xh, d, k, 0.0, True, return_fallback=True)
    torch.cuda.synchronize()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        simknn.build_knn_normalized(xf, xh, d, k, 0.0, True)      # warm-up on the side stream (allocations, function attributes)
    torch.cuda.current_stream().wait_stream(s)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        captured = simknn.build_knn_normalized(xf, xh, d, k, 0.0, True, return_fallback=True)
    g.replay()
    torch.cuda.synchronize()
    for a, b in zip(eager[:3], captured[:3]):
        assert torch.equal(a, b)
    assert int(captured[3][1]) == int(eager[3][1]) > 0            # the same rows were flagged for the retry pass
