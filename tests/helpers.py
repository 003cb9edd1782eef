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


def mdns_case(seed, n_way=2, k_shot=5, N=512, grid_xyz=True, noisy=(1,)):
    """A small support set for direct tests of the noise-suppression internals: per-class feature
    centres + noise (shots listed in `noisy` show another class), xyz on a coarse binary grid so that
    many foreground points lie EXACTLY on the faces between the cells of scale (2,2,1) and on the
    bounding box — the inclusive-bounds corner of the reference (models/mpti.py:355-367).
    Returns support_x (n_way, k_shot, 9, N), support_y (n_way, k_shot, N) int32,
    support_feat (n_way, k_shot, 192, N)."""
    import torch
    g = torch.Generator().manual_seed(seed)
    centres = torch.randn((n_way + 2, 192), generator=g)
    sx = torch.rand((n_way, k_shot, 9, N), generator=g)
    if grid_xyz:
        sx[:, :, :3] = torch.randint(0, 9, (n_way, k_shot, 3, N), generator=g).float() * 0.125
    sy = (torch.rand((n_way, k_shot, N), generator=g) < 0.3).to(torch.int32)
    sy[:, :, 0] = 1
    sf = torch.empty((n_way, k_shot, 192, N))
    for w in range(n_way):
        for k in range(k_shot):
            c = centres[n_way + (k % 2)] if k in noisy else centres[w]
            sf[w, k] = (c[:, None] + 0.6 * torch.randn((192, N), generator=g))
    return sx, sy, sf
