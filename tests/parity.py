"""Index-parity bookkeeping shared by the GPU tests, smoke() and bench.py (SURVEY.md §8(d) "parity gate").

A neighbour list "matches" the oracle when it is identical; it is "in band" when it differs only by
candidates whose FP64 scores are within `band` of each other / of the cut (summation-order noise of two FP32
implementations); anything else is out of band and fails the gate."""
import torch


def compare_lists(idx, cnt, idx_ref, cnt_ref, score64, thr, band=1e-6):
    """idx/idx_ref [R,k] (rank order, -1 padded), cnt/cnt_ref [R]; score64(row_ids, col_ids)->float64 scores.
    Returns dict(exact=, in_band=, out_of_band=, rows=[...out of band row ids])."""
    idx, cnt, idx_ref, cnt_ref = idx.long().cpu(), cnt.long().cpu(), idx_ref.long().cpu(), cnt_ref.long().cpu()
    R, k = idx.shape
    keep = torch.arange(k)[None, :] < cnt[:, None]
    keep_ref = torch.arange(k)[None, :] < cnt_ref[:, None]
    same = ((idx == idx_ref) | (~keep & ~keep_ref)).all(1) & (cnt == cnt_ref)
    bad_rows, in_band = [], 0
    for r in (~same).nonzero().flatten().tolist():
        a = idx[r, :cnt[r]]
        b = idx_ref[r, :cnt_ref[r]]
        rows = torch.full((max(a.numel(), b.numel(), 1),), r, dtype=torch.long)
        va = score64(rows[:a.numel()], a).double() if a.numel() else torch.zeros(0, dtype=torch.float64)
        vb = score64(rows[:b.numel()], b).double() if b.numel() else torch.zeros(0, dtype=torch.float64)
        m = min(a.numel(), b.numel())
        ok = bool((va[:m] - vb[:m]).abs().max() < band) if m else True
        # extra entries on either side must sit on the thr cut
        if a.numel() > m:
            ok = ok and bool((va[m:] - thr).abs().max() < band)
        if b.numel() > m:
            ok = ok and bool((vb[m:] - thr).abs().max() < band)
        if ok:
            in_band += 1
        else:
            bad_rows.append(r)
    return dict(exact=int(same.sum()), in_band=in_band, out_of_band=len(bad_rows), rows=bad_rows[:10])


def check_tie_order(idx, sim, cnt):
    """Exact FP32 ties inside a list must be ordered by ascending index (the documented tie-break)."""
    idx, sim, cnt = idx.long().cpu(), sim.cpu(), cnt.long().cpu()
    k = idx.size(1)
    if k < 2:
        return True
    valid = torch.arange(1, k)[None, :] < cnt[:, None]
    tie = (sim[:, 1:] == sim[:, :-1]) & valid
    return bool(((idx[:, 1:] > idx[:, :-1]) | ~tie).all()) and bool(((sim[:, 1:] <= sim[:, :-1]) | ~valid).all())
