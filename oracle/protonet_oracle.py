"""TEST INFRASTRUCTURE — CPU restatement of the reference's ProtoNet+MDNS episode
(reference models/protonet.py:780-858, ProtoNet_Contrast.forward with train=False).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this module; the
product path (r3dfsseg_b200) never does.  Pinned to the reference: tests/test_oracle_golden.py checks
it against tests/golden/golden_protonet.pt, which oracle/make_golden.py wrote by running the
reference's own module (under oracle/ref_shims.py) on the seeded episodes.
"""
from __future__ import annotations

from typing import Dict

import torch
import torch.nn.functional as F

from .mpti_oracle import get_features, mdns_multi_scale


def masked_features(feat: torch.Tensor, mask: torch.Tensor) -> torch.Tensor:
    """models/protonet.py:878-890 — feat (n_way, k_shot, D, N), mask (n_way, k_shot, N)."""
    mask = mask.unsqueeze(2)
    return torch.sum(feat * mask, dim=3) / (mask.sum(dim=3) + 1e-5)


def prototypes(fg_feat, bg_feat, clean_flag=None):
    """models/protonet.py:892-915 -> (n_way + 1, D): background first, then one per way."""
    n_way, k_shot, _ = fg_feat.shape
    rows = [bg_feat.sum(dim=(0, 1)) / (n_way * k_shot)]
    for way in range(n_way):
        if clean_flag is not None:
            m = clean_flag[way].unsqueeze(-1)
            rows.append(torch.sum(fg_feat[way] * m, dim=0) / torch.sum(clean_flag[way]))
        else:
            rows.append(fg_feat[way].sum(dim=0) / k_shot)
    return torch.stack(rows, 0)


def forward_episode(sd: Dict[str, torch.Tensor], support_x, support_y, query_x, query_y,
                    dgcnn_k: int = 20, mdns: bool = True, support_feat=None,
                    query_feat=None) -> Dict[str, object]:
    """`support_feat` (n_way*k_shot, D, N) / `query_feat` (n_query, D, N), when given, replace the
    getFeatures calls (used to check the head on features produced elsewhere)."""
    n_way, k_shot = support_y.shape[:2]
    N = support_y.shape[-1]
    sx = support_x.reshape(n_way * k_shot, -1, N)
    if support_feat is None:
        support_feat = get_features(sx, sd, dgcnn_k)
    D = support_feat.shape[1]
    support_feat = support_feat.reshape(n_way, k_shot, D, N)
    if query_feat is None:
        query_feat = get_features(query_x, sd, dgcnn_k)
    clean = None
    if mdns:
        _, clean = mdns_multi_scale(support_feat, support_y,
                                    support_x.reshape(n_way, k_shot, -1, N))
    fg = masked_features(support_feat, support_y)
    bg = masked_features(support_feat, torch.logical_not(support_y))
    protos = prototypes(fg, bg, clean)
    # models/protonet.py:917-940, method 'cosine', scaler 10
    sim = [F.cosine_similarity(query_feat, p[None, :, None], dim=1) * 10 for p in protos]
    query_pred = torch.stack(sim, dim=1)
    loss = F.cross_entropy(query_pred, query_y)
    return dict(query_pred=query_pred, loss=loss, clean_flag=clean, prototypes=protos)
