"""Tie-aware comparators shared by the parity tests (SURVEY.md §8c acceptance rules)."""
import torch


def knn_sets_match(idx_a, idx_b, scores, largest, rtol=1e-5):
    """Rows are compared as SETS; an index present on one side only is accepted when its score is
    within rtol * max(1, |s_k|) of the k-th best score (a tie the two sides broke differently).
    idx_*: (..., n, k) ; scores: (..., n, n) float64 ranking key.  Returns (ok, n_bad_rows)."""
    k = idx_a.shape[-1]
    a = idx_a.reshape(-1, k).long().cpu()
    b = idx_b.reshape(-1, k).long().cpu()
    s = scores.reshape(-1, scores.shape[-1]).double().cpu()
    sa, sb = a.sort(1)[0], b.sort(1)[0]
    rows = torch.nonzero((sa != sb).any(1)).flatten().tolist()
    bad = 0
    for r in rows:
        A, B = set(a[r].tolist()), set(b[r].tolist())
        kth = s[r, b[r]].min() if largest else s[r, b[r]].max()
        tol = rtol * max(1.0, abs(float(kth)))
        for j in (A ^ B):
            if abs(float(s[r, j]) - float(kth)) > tol:
                bad += 1
                break
    return bad == 0, bad
